// Spectral-clustering back end for sm_100a.  Replaces the arithmetic of
// SpectralCluster.__call__ (speakerlab/process/cluster.py:35-112):
//
//   get_sim_mat   :59-62   cosine affinity  = normalize(X) normalize(X)^T
//   p_pruning     :64-77   per row keep the `keep` largest entries
//   sym + get_laplacian :46,:79-84   M = 0.5 (P + P^T), zero diagonal, L = diag(sum |M|) - M
//   get_spec_embs :86-100  k smallest eigenpairs (scipy eigsh which='SM' in the reference)
//   cluster_embs  :102-105 k-means (Lloyd iterations here; k-means++ seeding stays on the host
//                          RNG exactly like sklearn, see b200spk/cluster.py)
//
// The affinity is a tensor-core GEMM at fp32 accuracy: every normalised row is split into three
// bf16 terms (x = hi + mid + lo) and ONE tcgen05 GEMM with K = 6 D accumulates the six
// significant cross terms in the fp32 TMEM accumulator (conv_gemm.cu).  fp32-level accuracy is
// needed because pruning keeps an exact count per row and near-ties decide which edges survive.
// Pruning is a per-row radix select in shared memory; the eigensolver is Lanczos with full
// re-orthogonalisation on sigma*I - L (sigma = Gershgorin bound), whose largest Ritz pairs are
// the smallest eigenpairs of L; the small tridiagonal problem is solved on the host (implicit QL).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cooperative_groups.h>

#include "ops.cuh"

namespace cg = cooperative_groups;

namespace spk {
namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ float block_sum(float v, float *sh) {      // sh: >= 32 floats
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (warp == 0) r = warp_sum(r);
    if (threadIdx.x == 0) sh[0] = r;
    __syncthreads();
    r = sh[0];
    return r;
}

// ---- row L2-normalise + 3-term bf16 split, laid out for the K = 6*Dp product
// A' = [hi hi mid hi lo mid],  B' = [hi mid hi lo hi mid]  ->  A' B'^T = sum of the six terms
__global__ void __launch_bounds__(256)
normalize_split_kernel(const float *__restrict__ X, int N, int D, int Np, int Dp, bf16 *__restrict__ A, bf16 *__restrict__ B,
                       float *__restrict__ Xn) {
    __shared__ float sh[32];
    const int row = blockIdx.x;
    float ss = 0.f;
    if (row < N)
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            const float v = X[(size_t)row * D + c];
            ss += v * v;
        }
    ss = block_sum(ss, sh);
    float inv = 0.f;
    if (row < N) inv = ss > 0.f ? 1.f / sqrtf(ss) : 1.f;        // sklearn normalize: zero rows stay zero
    for (int c = threadIdx.x; c < Dp; c += blockDim.x) {
        const float v = (row < N && c < D) ? X[(size_t)row * D + c] * inv : 0.f;
        if (Xn != nullptr) Xn[(size_t)row * Dp + c] = v;
        if (A != nullptr) {
            const bf16 hi = __float2bfloat16_rn(v);
            const float r1 = v - __bfloat162float(hi);
            const bf16 mid = __float2bfloat16_rn(r1);
            const bf16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
            bf16 *a = A + (size_t)row * 6 * Dp, *b = B + (size_t)row * 6 * Dp;
            a[c] = hi;          b[c] = hi;
            a[Dp + c] = hi;     b[Dp + c] = mid;
            a[2 * Dp + c] = mid; b[2 * Dp + c] = hi;
            a[3 * Dp + c] = hi;  b[3 * Dp + c] = lo;
            a[4 * Dp + c] = lo;  b[4 * Dp + c] = hi;
            a[5 * Dp + c] = mid; b[5 * Dp + c] = mid;
        }
    }
}

// ---- p-pruning: keep the `keep` largest entries of each row (ties: higher column index wins,
// i.e. what a stable ascending argsort zeroes first).  One CTA per row, row in shared memory,
// 4-pass 8-bit radix select on the order-preserving integer image of the floats.
__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
prune_rows_kernel(float *__restrict__ S, int N, int ld, int keep) {
    extern __shared__ unsigned shm[];
    unsigned *keys = shm;                 // [N]
    unsigned *hist = shm + N;             // [256]
    __shared__ unsigned s_prefix, s_mask;
    __shared__ int s_need;
    const int row = blockIdx.x;
    float *srow = S + (size_t)row * ld;
    for (int j = threadIdx.x; j < N; j += blockDim.x) keys[j] = f2key(srow[j]);
    if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_need = keep; }
    __syncthreads();
    // find the key of the keep-th largest element
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        const int shift = pass * 8;
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            const unsigned k = keys[j];
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int need = s_need;
            int b = 255;
            for (; b > 0; --b) {
                const int c = (int)hist[b];
                if (c >= need) break;
                need -= c;
            }
            s_need = need;
            s_prefix = prefix | ((unsigned)b << shift);
            s_mask = mask | (255u << shift);
        }
        __syncthreads();
    }
    const unsigned tau = s_prefix;       // key of the keep-th largest
    const int ties_to_keep = s_need;     // how many elements equal to tau survive
    // ties: keep the ones with the highest column index.  Count ties from the right.
    // (rare path; a serial scan by one thread is fine for the handful of rows that have ties)
    __shared__ int s_tie_cut;            // smallest column index among kept ties
    if (threadIdx.x == 0) {
        int left = ties_to_keep, cut = N;
        for (int j = N - 1; j >= 0 && left > 0; --j)
            if (keys[j] == tau) { cut = j; --left; }
        s_tie_cut = cut;
    }
    __syncthreads();
    const int cut = s_tie_cut;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const unsigned k = keys[j];
        const bool kept = (k > tau) || (k == tau && j >= cut);
        if (!kept) srow[j] = 0.f;
    }
}

// ---- L offdiag = -0.5 (P + P^T); diagonal filled by laplacian_diag_kernel
__global__ void __launch_bounds__(256)
symmetrize_neg_kernel(const float *__restrict__ P, float *__restrict__ L, int N, int ldp, int ldl) {
    __shared__ float t[32][33];
    const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = bj + r, j = bi + tx;                       // transposed block P[bj.., bi..]
        t[r][tx] = (i < N && j < N) ? P[(size_t)i * ldp + j] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = bi + r, j = bj + tx;
        if (i < N && j < N) {
            const float v = 0.5f * (P[(size_t)i * ldp + j] + t[tx][r]);
            L[(size_t)i * ldl + j] = (i == j) ? 0.f : -v;
        }
    }
}
__global__ void __launch_bounds__(256)
laplacian_diag_kernel(float *__restrict__ L, int N, int ld) {
    __shared__ float sh[32];
    const int row = blockIdx.x;
    float s = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x)
        if (j != row) s += fabsf(L[(size_t)row * ld + j]);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) L[(size_t)row * ld + row] = s;
}

// ---- Lanczos on sigma*I - L with full re-orthogonalisation, as ONE cooperative kernel.
//
// The pruned, symmetrised Laplacian has ~2*keep non-zeros per row (420 k of 23 M entries for a 1-hour meeting), so
// it is compacted once to CSR (3.4 MB: stays in L2) and every Lanczos step is a sparse mat-vec.  The whole step
// sequence runs inside one cooperative launch: CTA c owns a contiguous slice of rows, i.e. rows of the mat-vec AND
// the same rows of every basis vector (stored transposed, Vt[row][step], so a CTA only ever touches its own slice
// of the basis).  Per step three grid-wide barriers replace the eight kernel launches of the straightforward
// formulation:
//   1. w = sigma v_j - L v_j on the slice, partial dots  V^T w over the slice          -> barrier
//   2. h1 = sum of the partials (fixed CTA order), w1 = w - V h1, partials of V^T w1 and |w1|^2 -> barrier
//   3. h2 likewise, w2 = w1 - V h2 (classical Gram-Schmidt applied twice), beta^2 = |w1|^2 - |h2|^2,
//      v_{j+1} = w2 / beta, alpha_j = h1[j] + h2[j]                                     -> barrier
// All reductions have a fixed order, so the tridiagonal - and the labels downstream - are bit-reproducible.
constexpr int kCsrPerRow = 256;          // CSR capacity per row on average; denser matrices use the dense rows
constexpr int kLzThreads = 1024;     // 32 warps: one warp per row of the slice in the mat-vec and Gram-Schmidt passes

__global__ void __launch_bounds__(256)
csr_count_kernel(const float *__restrict__ L, int N, int ld, int *__restrict__ rowcnt) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const float *lr = L + (size_t)row * ld;
    int c = 0;
    for (int j = lane; j < N; j += 32) c += (j != row && lr[j] != 0.f) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) rowcnt[row] = c;
}
// exclusive scan of the row counts (one CTA; N is a few thousand) -> rowptr[N+1]; flags[0] = 1 when the CSR fits
__global__ void __launch_bounds__(1024)
csr_scan_kernel(const int *__restrict__ rowcnt, int N, int *__restrict__ rowptr, long long cap, int *__restrict__ flags) {
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < N; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < N ? rowcnt[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const int t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < N) rowptr[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rowptr[N] = carry;
        flags[0] = (long long)carry <= cap ? 1 : 0;
    }
}
__global__ void __launch_bounds__(256)
csr_fill_kernel(const float *__restrict__ L, int N, int ld, const int *__restrict__ rowptr, const int *__restrict__ flags,
                int *__restrict__ cols, float *__restrict__ vals, float *__restrict__ diag) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const float *lr = L + (size_t)row * ld;
    if (lane == 0) diag[row] = lr[row];
    if (!flags[0]) return;
    int base = rowptr[row];
    for (int j0 = 0; j0 < N; j0 += 32) {                 // ascending column order: deterministic layout
        const int j = j0 + lane;
        const float v = j < N ? lr[j] : 0.f;
        const bool nz = j < N && j != row && v != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (nz) {
            const int p = base + __popc(m & ((1u << lane) - 1u));
            cols[p] = j;
            vals[p] = v;
        }
        base += __popc(m);
    }
}

struct LzArgs {
    const float *L;          // dense Laplacian (row pitch ld), used when the CSR does not fit
    int N, ld;
    const int *rowptr, *cols, *flags;
    const float *vals, *diag;
    float *Vt;               // [N][pitch] basis, transposed: Vt[i][j] = v_j[i]
    int pitch;
    float *vbuf;             // [2][N] the current / next Lanczos vector, contiguous for the mat-vec gathers
    float *partial;          // [2][grid][pm] per-CTA partial reductions
    int pm;
    float *alpha, *beta, *dsigma;
    int m_from, m_to, rows_per_cta;
    int spitch;              // > 0: the CTA's rows of the basis live in shared memory with this pitch
};

// sum over the CTAs of partial[c][jj] for jj < n -> h_s[jj]: eight lanes per jj, each a fixed stride-8 walk over
// the CTAs, folded by a fixed shuffle tree (order independent of scheduling)
__device__ __forceinline__ void lz_reduce_partials(const float *part, int pm, int G, int n, float *h_s) {
    const int tid = threadIdx.x, sub = tid & 7;
    for (int jj0 = 0; jj0 < n; jj0 += kLzThreads / 8) {
        const int jj = jj0 + (tid >> 3);
        float acc = 0.f;
        if (jj < n)
            for (int c = sub; c < G; c += 8) acc += part[(size_t)c * pm + jj];
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (jj < n && sub == 0) h_s[jj] = acc;
    }
}
// partial[jj] = sum over the slice rows of vt[il][jj] * w_s[il] for jj < n: four lanes per jj
__device__ __forceinline__ void lz_partial_dots(const float *vt, int vpitch, int nr, const float *w_s, int n, float *out) {
    const int tid = threadIdx.x, sub = tid & 3;
    for (int jj0 = 0; jj0 < n; jj0 += kLzThreads / 4) {
        const int jj = jj0 + (tid >> 2);
        float acc = 0.f;
        if (jj < n)
            for (int il = sub; il < nr; il += 4) acc = fmaf(vt[(size_t)il * vpitch + jj], w_s[il], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (jj < n && sub == 0) out[jj] = acc;
    }
}

__global__ void __launch_bounds__(kLzThreads)
lanczos_coop_kernel(const LzArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float lz_sh[];
    const int R = a.rows_per_cta;
    float *w_s = lz_sh;                  // [R]
    float *h_s = lz_sh + R;              // [m_to + 2]
    float *red = h_s + a.m_to + 2;       // [32]
    float *vt_s = red + 32;              // [R][spitch] this CTA's rows of the basis, when they fit shared memory
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, G = gridDim.x, cta = blockIdx.x;
    const int r0 = min(a.N, cta * R), r1 = min(a.N, r0 + R), nr = r1 - r0;
    const bool csr = a.flags[0] != 0;
    float *part0 = a.partial, *part1 = a.partial + (size_t)G * a.pm;
    // the slice of the basis this CTA works on: shared memory copy (spitch > 0) or the global rows
    const int vpitch = a.spitch > 0 ? a.spitch : a.pitch;
    float *vt = a.spitch > 0 ? vt_s : a.Vt + (size_t)r0 * a.pitch;
    float sigma;

    if (a.m_from == 0) {
        // Gershgorin shift (unnormalised Laplacian: row sum of |offdiag| equals the diagonal) and the start vector:
        // a deterministic hash of the row index, normalised
        float mx = 0.f, ss = 0.f;
        for (int il = tid; il < nr; il += kLzThreads) {
            const int i = r0 + il;
            mx = fmaxf(mx, a.diag[i]);
            unsigned x = (unsigned)i * 2654435761u ^ 0x9E3779B9u;
            x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
            const float v = ((x >> 8) * (1.0f / 16777216.0f)) - 0.5f;
            w_s[il] = v;
            ss = fmaf(v, v, ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        if (tid == 0) {
            float m2 = 0.f;
            for (int w = 0; w < kLzThreads / 32; ++w) m2 = fmaxf(m2, red[w]);
            part0[(size_t)cta * a.pm] = m2;
        }
        ss = block_sum(ss, red);
        if (tid == 0) part0[(size_t)cta * a.pm + 1] = ss;
        grid.sync();
        float m2 = 0.f, tot = 0.f;
        for (int c = 0; c < G; ++c) {
            m2 = fmaxf(m2, part0[(size_t)c * a.pm]);
            tot += part0[(size_t)c * a.pm + 1];
        }
        sigma = 2.f * m2 + 1e-3f;
        const float inv = rsqrtf(tot);
        for (int il = tid; il < nr; il += kLzThreads) {
            const float v = w_s[il] * inv;
            a.Vt[(size_t)(r0 + il) * a.pitch] = v;
            if (a.spitch > 0) vt_s[(size_t)il * a.spitch] = v;
            a.vbuf[r0 + il] = v;
        }
        if (cta == 0 && tid == 0) a.dsigma[0] = sigma;
        grid.sync();
    } else {
        sigma = a.dsigma[0];
        if (a.spitch > 0) {             // continue a run: pull this slice of the basis built so far into shared memory
            for (int idx = tid; idx < nr * (a.m_from + 1); idx += kLzThreads) {
                const int il = idx / (a.m_from + 1), jj = idx - il * (a.m_from + 1);
                vt_s[(size_t)il * a.spitch + jj] = a.Vt[(size_t)(r0 + il) * a.pitch + jj];
            }
            __syncthreads();
        }
    }

    for (int j = a.m_from; j < a.m_to; ++j) {
        const float *v = a.vbuf + (size_t)(j & 1) * a.N;
        float *vnext = a.vbuf + (size_t)((j + 1) & 1) * a.N;
        // ---- phase 1: w = sigma v - L v on the slice (one warp per row), partial V^T w
        for (int il = warp; il < nr; il += kLzThreads / 32) {
            const int i = r0 + il;
            float s = 0.f;
            if (csr) {
                const int p1 = a.rowptr[i + 1];
                for (int p = a.rowptr[i] + lane; p < p1; p += 32) s = fmaf(a.vals[p], v[a.cols[p]], s);
            } else {
                const float *lr = a.L + (size_t)i * a.ld;
                for (int c = lane; c < a.N; c += 32) s = fmaf(c != i ? lr[c] : 0.f, v[c], s);
            }
            s = warp_sum(s);
            if (lane == 0) w_s[il] = (sigma - a.diag[i]) * v[i] - s;
        }
        __syncthreads();
        lz_partial_dots(vt, vpitch, nr, w_s, j + 1, part0 + (size_t)cta * a.pm);
        grid.sync();
        // ---- phase 2: h1, first Gram-Schmidt pass, partials of the second
        lz_reduce_partials(part0, a.pm, G, j + 1, h_s);
        __syncthreads();
        const float h1j = h_s[j];
        for (int il = warp; il < nr; il += kLzThreads / 32) {
            const float *vr = vt + (size_t)il * vpitch;
            float s = 0.f;
            for (int jj = lane; jj <= j; jj += 32) s = fmaf(h_s[jj], vr[jj], s);
            s = warp_sum(s);
            if (lane == 0) w_s[il] -= s;
        }
        __syncthreads();
        lz_partial_dots(vt, vpitch, nr, w_s, j + 1, part1 + (size_t)cta * a.pm);
        {
            float ss = 0.f;
            for (int il = tid; il < nr; il += kLzThreads) ss = fmaf(w_s[il], w_s[il], ss);
            ss = block_sum(ss, red);
            if (tid == 0) part1[(size_t)cta * a.pm + j + 1] = ss;
        }
        grid.sync();
        // ---- phase 3: h2, second pass, normalise -> v_{j+1}
        lz_reduce_partials(part1, a.pm, G, j + 2, h_s);
        __syncthreads();
        float hh = 0.f;
        for (int jj = tid; jj <= j; jj += kLzThreads) hh = fmaf(h_s[jj], h_s[jj], hh);
        hh = block_sum(hh, red);
        const float beta2 = fmaxf(h_s[j + 1] - hh, 0.f);      // |w1 - V h2|^2 for an orthonormal basis
        const float b = sqrtf(beta2);
        const float inv = b > 0.f ? 1.f / b : 0.f;
        for (int il = warp; il < nr; il += kLzThreads / 32) {
            float *vr = vt + (size_t)il * vpitch;
            float s = 0.f;
            for (int jj = lane; jj <= j; jj += 32) s = fmaf(h_s[jj], vr[jj], s);
            s = warp_sum(s);
            if (lane == 0) {
                const float vn = (w_s[il] - s) * inv;
                vr[j + 1] = vn;
                if (a.spitch > 0) a.Vt[(size_t)(r0 + il) * a.pitch + j + 1] = vn;
                vnext[r0 + il] = vn;
            }
        }
        if (cta == 0 && tid == 0) {
            a.alpha[j] = h1j + h_s[j];
            a.beta[j] = b;
        }
        grid.sync();
    }
}

// evecs[i, c] = sum_j Vt[i][j] * Sm[j, c]     (Ritz vectors; Sm is m x k row-major on device; one warp per row)
__global__ void __launch_bounds__(256)
ritz_kernel(const float *__restrict__ Vt, int N, int pitch, int m, const float *__restrict__ Sm, int k, float *__restrict__ out) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const float *vr = Vt + (size_t)i * pitch;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int j = lane; j < m; j += 32) {
        const float v = vr[j];
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (c < k) acc[c] = fmaf(v, Sm[j * k + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < 32; ++c)
        if (c < k) {
            const float r = warp_sum(acc[c]);
            if (lane == 0) out[(size_t)i * k + c] = r;
        }
}

// ---- k-means (Lloyd).  The update is a fixed-order tree reduction per (cluster, dimension), so
// the centres - and with them labels of boundary points - are bit-reproducible run to run.
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const float *__restrict__ P, int N, int d, int k, const float *__restrict__ C, int *__restrict__ labels,
                     float *__restrict__ inertia, int *__restrict__ changed) {
    extern __shared__ float cen[];        // [k*d]
    for (int i = threadIdx.x; i < k * d; i += blockDim.x) cen[i] = C[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float my_in = 0.f;
    if (i < N) {
        const float *p = P + (size_t)i * d;
        float best = 3.4e38f;
        int bj = 0;
        for (int j = 0; j < k; ++j) {
            float s = 0.f;
            for (int c = 0; c < d; ++c) {
                const float df = p[c] - cen[j * d + c];
                s = fmaf(df, df, s);
            }
            if (s < best) { best = s; bj = j; }      // first minimum wins, like np.argmin
        }
        if (labels[i] != bj) { labels[i] = bj; atomicAdd(changed, 1); }
        my_in = best;
    }
    my_in = warp_sum(my_in);
    if ((threadIdx.x & 31) == 0) atomicAdd(inertia, my_in);      // diagnostic only
}
// one CTA per cluster: new centre = mean of its points (empty clusters keep their centre)
__global__ void __launch_bounds__(256)
kmeans_update_kernel(const float *__restrict__ P, int N, int d, const int *__restrict__ labels, float *__restrict__ C,
                     float *__restrict__ shift2) {
    __shared__ float sh[32];
    const int j = blockIdx.x;
    float cnt = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) cnt += (labels[i] == j) ? 1.f : 0.f;
    cnt = block_sum(cnt, sh);
    if (cnt == 0.f) return;
    float sh2 = 0.f;
    for (int c = 0; c < d; ++c) {
        float s = 0.f;
        for (int i = threadIdx.x; i < N; i += blockDim.x)
            if (labels[i] == j) s += P[(size_t)i * d + c];
        s = block_sum(s, sh);
        const float nc = s / cnt;
        const float df = nc - C[j * d + c];
        sh2 = fmaf(df, df, sh2);
        __syncthreads();
        if (threadIdx.x == 0) C[j * d + c] = nc;
    }
    if (threadIdx.x == 0) atomicAdd(shift2, sh2);      // k addends: order-insensitive to ~1 ulp, only compared to tol
}

__global__ void cosine_pairs_kernel(const float *__restrict__ E, long long N, int D, const int *__restrict__ a,
                                    const int *__restrict__ b, long long n_pairs, float *__restrict__ out) {
    const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pair >= n_pairs) return;
    const float *x = E + (size_t)a[pair] * D, *y = E + (size_t)b[pair] * D;
    float xy = 0.f, xx = 0.f, yy = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float u = x[c], v = y[c];
        xy = fmaf(u, v, xy); xx = fmaf(u, u, xx); yy = fmaf(v, v, yy);
    }
    xy = warp_sum(xy); xx = warp_sum(xx); yy = warp_sum(yy);
    if (lane == 0) {
        const float nx = xx > 0.f ? sqrtf(xx) : 1.f, ny = yy > 0.f ? sqrtf(yy) : 1.f;   // sklearn normalize semantics
        out[pair] = xy / (nx * ny);
    }
}

// ---- host: symmetric tridiagonal eigen-decomposition (implicit QL with Wilkinson shifts)
// d[0..m) diagonal, e[0..m-1) off-diagonal; on return d = eigenvalues (unsorted), z = m x m
// eigenvectors (column c in z[r*m + c]).  Classic tql2 recurrence in double precision.
bool tridiag_eig(std::vector<double> &d, std::vector<double> &e, std::vector<double> &z, int m) {
    z.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) z[(size_t)i * m + i] = 1.0;
    e.resize(m);
    e[m - 1] = 0.0;
    for (int l = 0; l < m; ++l) {
        int iter = 0, mm;
        do {
            for (mm = l; mm < m - 1; ++mm) {
                const double dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
                if (std::fabs(e[mm]) <= 2.3e-16 * dd) break;
            }
            if (mm != l) {
                if (++iter > 200) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = mm - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    r = std::hypot(f, g);
                    e[i + 1] = r;
                    if (r == 0.0) {
                        d[i + 1] -= p;
                        e[mm] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - b;
                    for (int k = 0; k < m; ++k) {
                        f = z[(size_t)k * m + i + 1];
                        z[(size_t)k * m + i + 1] = s * z[(size_t)k * m + i] + c * f;
                        z[(size_t)k * m + i] = c * z[(size_t)k * m + i] - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[mm] = 0.0;
            }
        } while (mm != l);
    }
    return true;
}

struct AffinityPlan {
    int Np, Dp;
    bool tensor;
    int64_t off_a, off_b, off_xn, off_s, total;
};
AffinityPlan affinity_plan(int64_t N, int64_t D) {
    AffinityPlan p;
    p.Np = (int)align_up(N, 16);
    p.Dp = (int)align_up(D, 16);
    p.tensor = p.Np >= 128;
    int64_t cur = 0;
    p.off_a = cur; cur += align_up((int64_t)p.Np * 6 * p.Dp * 2, 1024);
    p.off_b = cur; cur += align_up((int64_t)p.Np * 6 * p.Dp * 2, 1024);
    p.off_xn = cur; cur += align_up((int64_t)p.Np * p.Dp * 4, 1024);
    p.off_s = cur; cur += align_up((int64_t)p.Np * p.Np * 4, 1024);
    p.total = cur;
    return p;
}

// normalised rows -> cosine similarity S [Np x Np] (row pitch Np) in the workspace; shared by the spectral and AHC paths
int cosine_matrix(const float *X, int64_t N, int64_t D, const AffinityPlan &p, char *ws, cudaStream_t s) {
    bf16 *A = reinterpret_cast<bf16 *>(ws + p.off_a), *B = reinterpret_cast<bf16 *>(ws + p.off_b);
    float *Xn = reinterpret_cast<float *>(ws + p.off_xn);
    const int Np = p.Np, Dp = p.Dp;
    float *S = reinterpret_cast<float *>(ws + p.off_s);
    normalize_split_kernel<<<Np, 256, 0, s>>>(X, (int)N, (int)D, Np, Dp, p.tensor ? A : nullptr, p.tensor ? B : nullptr, Xn);
    int rc = check_launch("normalize_split_kernel");
    if (rc != SPK_OK) return rc;
    ConvArgs a{};
    a.y = S; a.B = 1; a.H = 1; a.W = Np; a.Ho = 1; a.Wo = Np; a.Cout = Np;
    a.KH = a.KW = a.sh = a.sw = a.dh = a.dw = 1;
    a.out_ld = Np; a.gate_win = 1; a.gate_nwin = 1; a.M = Np;
    if (p.tensor) {
        a.x = A; a.w = B; a.Cin = 6 * Dp; a.K = 6 * Dp; a.in_ld = 6 * Dp;
        if (!conv_gemm_supported(a, SPK_DT_BF16)) {
            set_error("affinity: GEMM shape not supported (N=%d, D=%d)", Np, Dp);
            return SPK_ERR_UNSUPPORTED;
        }
        return launch_conv_gemm(a, SPK_DT_F32, SPK_DT_F32, s);
    }
    a.x = Xn; a.w = Xn; a.Cin = Dp; a.K = Dp; a.in_ld = Dp;
    return launch_conv_simt(a, SPK_DT_F32, SPK_DT_F32, SPK_DT_F32, s);
}

// ------------------------------------------------------------------ agglomerative clustering (average linkage)
// speakerlab/process/cluster.py:139-156: average-linkage AHC on the distance -cos, cut where the linkage distance
// exceeds -fix_cos_thr.  UPGMA is monotone, so the flat clustering is "keep merging the closest pair while its
// distance is <= the cut"; no dendrogram is stored.  The distance matrix lives in HBM/L2 (it IS the negated affinity);
// every row keeps its nearest active neighbour, one persistent CTA runs the merge loop:
//   argmin over rows -> Lance-Williams update of the merged row and column -> refresh the neighbours it invalidated.
constexpr int kAhcThreads = 1024;
constexpr float kAhcInf = 3.0e38f;

__global__ void __launch_bounds__(256)
ahc_init_kernel(float *Dm, int N, int ld, int *nn, float *nnd) {
    __shared__ float s_v[256];
    __shared__ int s_i[256];
    const int i = blockIdx.x;
    float *row = Dm + (size_t)i * ld;
    float best = kAhcInf;
    int bj = -1;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float d = (j == i) ? kAhcInf : -row[j];
        row[j] = d;
        if (d < best) { best = d; bj = j; }
    }
    s_v[threadIdx.x] = best; s_i[threadIdx.x] = bj;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const float v = s_v[threadIdx.x + o];
            const int k = s_i[threadIdx.x + o];
            if (v < s_v[threadIdx.x] || (v == s_v[threadIdx.x] && k >= 0 && (s_i[threadIdx.x] < 0 || k < s_i[threadIdx.x]))) {
                s_v[threadIdx.x] = v; s_i[threadIdx.x] = k;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { nn[i] = s_i[0]; nnd[i] = s_v[0]; }
}

// block-wide (value, index) argmin with ties to the smaller index; result broadcast through shared memory
__device__ __forceinline__ void block_argmin(float &v, int &idx, float *s_v, int *s_i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi >= 0 && (idx < 0 || oi < idx))) { v = ov; idx = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { s_v[warp] = v; s_i[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? s_v[lane] : kAhcInf;
        idx = lane < (int)(blockDim.x >> 5) ? s_i[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov < v || (ov == v && oi >= 0 && (idx < 0 || oi < idx))) { v = ov; idx = oi; }
        }
        if (lane == 0) { s_v[0] = v; s_i[0] = idx; }
    }
    __syncthreads();
    v = s_v[0]; idx = s_i[0];
    __syncthreads();
}

__global__ void __launch_bounds__(kAhcThreads, 1)
ahc_merge_kernel(float *Dm, int N, int ld, int *nn, float *nnd, int *size, int *parent, float cut, int *labels, int *n_clusters) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    __shared__ int s_todo[kAhcThreads];
    __shared__ int s_ntodo;
    for (int k = threadIdx.x; k < N; k += blockDim.x) { size[k] = 1; parent[k] = k; }
    __syncthreads();
    for (int step = 0; step < N - 1; ++step) {
        // ---- closest pair
        float v = kAhcInf;
        int i = -1;
        for (int k = threadIdx.x; k < N; k += blockDim.x)
            if (size[k] > 0 && (nnd[k] < v || (nnd[k] == v && (i < 0 || k < i)))) { v = nnd[k]; i = k; }
        block_argmin(v, i, s_v, s_i);
        if (i < 0 || !(v <= cut)) break;
        const int j = nn[i];
        const float ni = (float)size[i], nj = (float)size[j], inv = 1.f / (ni + nj);
        if (threadIdx.x == 0) s_ntodo = 0;
        __syncthreads();
        // ---- Lance-Williams (average): d(i+j, k) = (ni d(i,k) + nj d(j,k)) / (ni + nj); row i becomes the merged cluster
        float *ri = Dm + (size_t)i * ld, *rj = Dm + (size_t)j * ld;
        float bv = kAhcInf;
        int bk = -1;
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            if (size[k] <= 0 || k == i || k == j) continue;
            const float d = (ni * ri[k] + nj * rj[k]) * inv;
            ri[k] = d;
            Dm[(size_t)k * ld + i] = d;
            Dm[(size_t)k * ld + j] = kAhcInf;
            if (d < bv || (d == bv && (bk < 0 || k < bk))) { bv = d; bk = k; }
            // neighbour bookkeeping of row k
            if (nn[k] == i || nn[k] == j) {
                const int slot = atomicAdd(&s_ntodo, 1);
                if (slot < kAhcThreads) s_todo[slot] = k;
            } else if (d < nnd[k] || (d == nnd[k] && i < nn[k])) {
                nn[k] = i; nnd[k] = d;
            }
        }
        block_argmin(bv, bk, s_v, s_i);
        if (threadIdx.x == 0) {
            ri[j] = kAhcInf; rj[i] = kAhcInf;
            nn[i] = bk; nnd[i] = bk >= 0 ? bv : kAhcInf;
            size[i] += size[j]; size[j] = 0; parent[j] = i;
            nnd[j] = kAhcInf; nn[j] = -1;
        }
        __syncthreads();
        // ---- rows whose nearest neighbour was i or j: full rescan (rows are distinct, any order gives the same result)
        const int ntodo = s_ntodo;
        if (ntodo > kAhcThreads) {          // cannot happen for N <= kAhcThreads * (N / kAhcThreads); rescan everything
            for (int k = 0; k < N; ++k) {
                if (size[k] <= 0 || k == i) continue;
                const float *rk = Dm + (size_t)k * ld;
                float tv = kAhcInf; int tk = -1;
                for (int q = threadIdx.x; q < N; q += blockDim.x)
                    if (size[q] > 0 && q != k && (rk[q] < tv || (rk[q] == tv && (tk < 0 || q < tk)))) { tv = rk[q]; tk = q; }
                block_argmin(tv, tk, s_v, s_i);
                if (threadIdx.x == 0) { nn[k] = tk; nnd[k] = tk >= 0 ? tv : kAhcInf; }
            }
        } else {
            for (int t = 0; t < ntodo; ++t) {
                const int k = s_todo[t];
                const float *rk = Dm + (size_t)k * ld;
                float tv = kAhcInf; int tk = -1;
                for (int q = threadIdx.x; q < N; q += blockDim.x)
                    if (size[q] > 0 && q != k && (rk[q] < tv || (rk[q] == tv && (tk < 0 || q < tk)))) { tv = rk[q]; tk = q; }
                block_argmin(tv, tk, s_v, s_i);
                if (threadIdx.x == 0) { nn[k] = tk; nnd[k] = tk >= 0 ? tv : kAhcInf; }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- flat labels: root of the merge chain, renumbered 0..K-1 in order of the smallest member index
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
        int r = k;
        while (parent[r] != r) r = parent[r];
        labels[k] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int next = 0;
        for (int k = 0; k < N; ++k)
            if (parent[k] == k) size[k] = -(++next);      // reuse size[] of the roots: -(new label + 1)
        *n_clusters = next;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) labels[k] = -size[labels[k]] - 1;
}

struct AhcPlan { AffinityPlan aff; int64_t off_nn, off_nnd, off_size, off_parent, off_k, total; };
AhcPlan ahc_plan(int64_t N, int64_t D) {
    AhcPlan p;
    p.aff = affinity_plan(N, D);
    int64_t cur = p.aff.total;
    p.off_nn = cur; cur += align_up(N * 4, 256);
    p.off_nnd = cur; cur += align_up(N * 4, 256);
    p.off_size = cur; cur += align_up(N * 4, 256);
    p.off_parent = cur; cur += align_up(N * 4, 256);
    p.off_k = cur; cur += 256;
    p.total = cur;
    return p;
}

}  // namespace
}  // namespace spk

using namespace spk;

extern "C" int64_t spk_affinity_workspace_bytes(int64_t N, int64_t D) {
    if (N <= 0 || D <= 0) return 0;
    return affinity_plan(N, D).total;
}

// L must hold Np x Np floats with Np = N rounded up to 16 (row pitch Np); rows/cols >= N are not written.
extern "C" int spk_affinity_laplacian(const float *X, int64_t N, int64_t D, int64_t keep, float *L, void *workspace,
                                      int64_t workspace_bytes, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    NvtxRange range("spk_affinity_laplacian");
    SPK_REQUIRE(X != nullptr && L != nullptr, "null buffer");
    SPK_REQUIRE(N >= 2 && D >= 1 && N < (1 << 24), "bad shape N=%lld D=%lld", (long long)N, (long long)D);
    SPK_REQUIRE(keep >= 1 && keep <= N, "keep=%lld out of range", (long long)keep);
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const AffinityPlan p = affinity_plan(N, D);
    if (workspace_bytes < p.total || workspace == nullptr) {
        set_error("workspace too small: need %lld bytes", (long long)p.total);
        return SPK_ERR_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    const int Np = p.Np;
    float *S = reinterpret_cast<float *>(ws + p.off_s);     // affinity, pruned in place
    rc = cosine_matrix(X, N, D, p, ws, s);
    if (rc != SPK_OK) return rc;
    const size_t sh = ((size_t)N + 256) * sizeof(unsigned);
    if (sh > 200 * 1024) {
        set_error("prune: N=%lld rows do not fit shared memory", (long long)N);
        return SPK_ERR_UNSUPPORTED;
    }
    static bool attr_done = false;
    if (!attr_done) {
        SPK_CUDA_OK(cudaFuncSetAttribute(prune_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    prune_rows_kernel<<<(unsigned)N, 256, sh, s>>>(S, (int)N, Np, (int)keep);
    rc = check_launch("prune_rows_kernel");
    if (rc != SPK_OK) return rc;
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((N + 31) / 32));
    symmetrize_neg_kernel<<<grid, 256, 0, s>>>(S, L, (int)N, Np, Np);
    rc = check_launch("symmetrize_neg_kernel");
    if (rc != SPK_OK) return rc;
    laplacian_diag_kernel<<<(unsigned)N, 256, 0, s>>>(L, (int)N, Np);
    return check_launch("laplacian_diag_kernel");
}

namespace {
struct EigPlan {
    int m_max, pitch, grid, rows_per_cta, pm;
    int64_t cap, off_V, off_vbuf, off_partial, off_alpha, off_beta, off_sm, off_sigma, off_rowcnt, off_rowptr, off_cols, off_vals,
        off_diag, off_flags, total;
};
EigPlan eig_plan(int64_t N, int k) {
    EigPlan p;
    // Krylov budget: clustered eigenvalues at the edge of the bulk (and exact multiplicities, which a
    // single-vector Lanczos only separates through round-off) can need several hundred steps
    p.m_max = (int)std::min<int64_t>(N, std::max(48 * k, 1024));
    p.pitch = (p.m_max + 1 + 31) & ~31;
    p.grid = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(), (N + 7) / 8));      // one CTA per SM: co-resident
    p.rows_per_cta = (int)((N + p.grid - 1) / p.grid);
    p.pm = p.m_max + 2;
    p.cap = std::min<int64_t>(N * N, (int64_t)kCsrPerRow * N);
    int64_t cur = 0;
    auto take = [&](int64_t bytes) { int64_t o = cur; cur += align_up(bytes, 256); return o; };
    p.off_V = take((int64_t)p.pitch * N * 4);
    p.off_vbuf = take(2 * N * 4);
    p.off_partial = take(2ll * p.grid * p.pm * 4);
    p.off_alpha = take((int64_t)(p.m_max + 1) * 4);
    p.off_beta = take((int64_t)(p.m_max + 1) * 4);
    p.off_sm = take((int64_t)p.m_max * 32 * 4);
    p.off_sigma = take(256);
    p.off_rowcnt = take(N * 4);
    p.off_rowptr = take((N + 1) * 4);
    p.off_cols = take(p.cap * 4);
    p.off_vals = take(p.cap * 4);
    p.off_diag = take(N * 4);
    p.off_flags = take(256);
    p.total = cur;
    return p;
}
}  // namespace

extern "C" int64_t spk_eig_workspace_bytes(int64_t N, int32_t k) {
    if (N <= 0 || k <= 0) return 0;
    return eig_plan(N, k).total;
}

namespace {
struct EigBufs { float *V, *vbuf, *partial, *alpha, *beta, *Sm, *dsigma, *vals, *diag; int *rowcnt, *rowptr, *cols, *flags; };
EigBufs eig_bufs(void *workspace, const EigPlan &p) {
    char *ws = static_cast<char *>(workspace);
    EigBufs b;
    b.V = reinterpret_cast<float *>(ws + p.off_V); b.vbuf = reinterpret_cast<float *>(ws + p.off_vbuf);
    b.partial = reinterpret_cast<float *>(ws + p.off_partial);
    b.alpha = reinterpret_cast<float *>(ws + p.off_alpha); b.beta = reinterpret_cast<float *>(ws + p.off_beta);
    b.Sm = reinterpret_cast<float *>(ws + p.off_sm); b.dsigma = reinterpret_cast<float *>(ws + p.off_sigma);
    b.vals = reinterpret_cast<float *>(ws + p.off_vals); b.diag = reinterpret_cast<float *>(ws + p.off_diag);
    b.rowcnt = reinterpret_cast<int *>(ws + p.off_rowcnt); b.rowptr = reinterpret_cast<int *>(ws + p.off_rowptr);
    b.cols = reinterpret_cast<int *>(ws + p.off_cols); b.flags = reinterpret_cast<int *>(ws + p.off_flags);
    return b;
}
}  // namespace

// Extend the Krylov basis kept in `workspace` from m_from to m_to vectors (m_from = 0 starts a new
// run: Gershgorin shift + deterministic start vector).  On return alpha_host[0..m_to) and
// beta_host[0..m_to) hold the Lanczos tridiagonal of sigma*I - L and *sigma_host the shift.
extern "C" int spk_lanczos_extend(const float *L, int64_t N, int32_t k, int32_t m_from, int32_t m_to, float *alpha_host,
                                  float *beta_host, float *sigma_host, void *workspace, int64_t workspace_bytes,
                                  void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    NvtxRange range("spk_lanczos_extend");
    SPK_REQUIRE(L != nullptr && alpha_host != nullptr && beta_host != nullptr && sigma_host != nullptr, "null buffer");
    SPK_REQUIRE(N >= 2 && k >= 1 && k <= 32, "bad N=%lld k=%d", (long long)N, k);
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const EigPlan p = eig_plan(N, k);
    if (workspace_bytes < p.total || workspace == nullptr) {
        set_error("workspace too small: need %lld bytes", (long long)p.total);
        return SPK_ERR_WORKSPACE;
    }
    SPK_REQUIRE(m_from >= 0 && m_to > m_from && m_to <= p.m_max, "bad Krylov range [%d,%d) (max %d)", m_from, m_to, p.m_max);
    const int n = (int)N, ld = (int)align_up(N, 16);
    const EigBufs b = eig_bufs(workspace, p);
    if (m_from == 0) {
        // compact the Laplacian once: row counts -> scan -> (column, value) runs in ascending column order
        csr_count_kernel<<<(n + 7) / 8, 256, 0, s>>>(L, n, ld, b.rowcnt);
        csr_scan_kernel<<<1, 1024, 0, s>>>(b.rowcnt, n, b.rowptr, (long long)p.cap, b.flags);
        csr_fill_kernel<<<(n + 7) / 8, 256, 0, s>>>(L, n, ld, b.rowptr, b.flags, b.cols, b.vals, b.diag);
        count_launch(3);
    }
    LzArgs a{};
    a.L = L; a.N = n; a.ld = ld; a.rowptr = b.rowptr; a.cols = b.cols; a.flags = b.flags; a.vals = b.vals; a.diag = b.diag;
    a.Vt = b.V; a.pitch = p.pitch; a.vbuf = b.vbuf; a.partial = b.partial; a.pm = p.pm;
    a.alpha = b.alpha; a.beta = b.beta; a.dsigma = b.dsigma; a.m_from = m_from; a.m_to = m_to; a.rows_per_cta = p.rows_per_cta;
    size_t sh = (size_t)(p.rows_per_cta + m_to + 2 + 32) * sizeof(float);
    SPK_REQUIRE(sh <= 64 * 1024, "lanczos: %lld rows per CTA do not fit shared memory", (long long)p.rows_per_cta);
    const int spitch = (m_to + 1) | 1;                     // odd pitch: the column walks of the dots stay conflict-free
    const size_t slice = (size_t)p.rows_per_cta * spitch * sizeof(float);
    a.spitch = sh + slice <= 200 * 1024 ? spitch : 0;      // else the basis rows are read from global memory (L2)
    if (a.spitch > 0) sh += slice;
    if (sh > 48 * 1024) SPK_CUDA_OK(cudaFuncSetAttribute(lanczos_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    void *kargs[] = {&a};
    SPK_CUDA_OK(cudaLaunchCooperativeKernel((const void *)lanczos_coop_kernel, dim3(p.grid), dim3(kLzThreads), kargs, sh, s));
    count_launch(1);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        set_error("lanczos launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
    SPK_CUDA_OK(cudaMemcpyAsync(sigma_host, b.dsigma, sizeof(float), cudaMemcpyDeviceToHost, s));
    SPK_CUDA_OK(cudaMemcpyAsync(alpha_host, b.alpha, m_to * sizeof(float), cudaMemcpyDeviceToHost, s));
    SPK_CUDA_OK(cudaMemcpyAsync(beta_host, b.beta, m_to * sizeof(float), cudaMemcpyDeviceToHost, s));
    SPK_CUDA_OK(cudaStreamSynchronize(s));
    return SPK_OK;
}

// Ritz vectors: evecs[N,k] = V[0..m)^T S, S = host [m,k] (eigenvectors of the tridiagonal, one per column)
extern "C" int spk_lanczos_ritz(int64_t N, int32_t k, int32_t m, const float *S_host, float *evecs, void *workspace,
                                int64_t workspace_bytes, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(S_host != nullptr && evecs != nullptr && workspace != nullptr, "null buffer");
    SPK_REQUIRE(N >= 2 && k >= 1 && k <= 32, "bad N=%lld k=%d", (long long)N, k);
    const EigPlan p = eig_plan(N, k);
    SPK_REQUIRE(workspace_bytes >= p.total && m >= 1 && m <= p.m_max, "bad workspace or m=%d", m);
    const EigBufs b = eig_bufs(workspace, p);
    SPK_CUDA_OK(cudaMemcpyAsync(b.Sm, S_host, (size_t)m * k * sizeof(float), cudaMemcpyHostToDevice, s));
    ritz_kernel<<<((int)N + 7) / 8, 256, 0, s>>>(b.V, (int)N, p.pitch, m, b.Sm, k, evecs);
    int rc = check_launch("ritz_kernel");
    if (rc != SPK_OK) return rc;
    SPK_CUDA_OK(cudaStreamSynchronize(s));      // S_host may be a temporary
    return SPK_OK;
}

extern "C" int32_t spk_lanczos_max_dim(int64_t N, int32_t k) { return (N > 0 && k > 0) ? eig_plan(N, k).m_max : 0; }

// Self-contained variant for C callers: same Lanczos run, tridiagonal solved by the built-in
// implicit-QL routine (O(m^3) with vectors - the Python mirror uses LAPACK's O(m k) instead).
extern "C" int spk_eig_smallest(const float *L, int64_t N, int32_t k, float *evals_host, float *evecs, void *workspace,
                                int64_t workspace_bytes, void *stream_) {
    SPK_REQUIRE(L != nullptr && evals_host != nullptr && evecs != nullptr, "null buffer");
    SPK_REQUIRE(N >= 2 && k >= 1 && k <= 32 && k <= N, "bad N=%lld k=%d (k <= 32)", (long long)N, k);
    const EigPlan p = eig_plan(N, k);
    std::vector<float> ha(p.m_max + 1), hb(p.m_max + 1);
    std::vector<double> d, e, z;
    std::vector<int> order;
    float sigma = 0.f;
    int m_done = 0, m_target = std::min(p.m_max, std::max(8 * k, 96));
    for (;;) {
        int rc = spk_lanczos_extend(L, N, k, m_done, m_target, ha.data(), hb.data(), &sigma, workspace, workspace_bytes, stream_);
        if (rc != SPK_OK) return rc;
        m_done = m_target;
        const int m = m_done;
        d.assign(m, 0.0);
        e.assign(m, 0.0);
        for (int i = 0; i < m; ++i) d[i] = ha[i];
        for (int i = 0; i + 1 < m; ++i) e[i] = hb[i];
        if (!tridiag_eig(d, e, z, m)) {
            set_error("tridiagonal QL did not converge");
            return SPK_ERR_KERNEL;
        }
        order.resize(m);
        for (int i = 0; i < m; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return d[x] > d[y]; });   // largest of sigma*I - L
        double worst = 0.0;      // residual estimate of Ritz pair c: |beta_m * z[m-1][c]|
        for (int c = 0; c < k; ++c) worst = std::max(worst, std::fabs((double)hb[m - 1] * z[(size_t)(m - 1) * m + order[c]]));
        if (worst <= 1e-4 || m_done >= p.m_max || m_done >= (int)N) break;
        m_target = std::min(p.m_max, m_done + std::max(6 * k, 96));
    }
    const int m = m_done;
    std::vector<float> sm((size_t)m * k);
    for (int c = 0; c < k; ++c) {
        evals_host[c] = (float)((double)sigma - d[order[c]]);
        for (int j = 0; j < m; ++j) sm[(size_t)j * k + c] = (float)z[(size_t)j * m + order[c]];
    }
    int rc = spk_lanczos_ritz(N, k, m, sm.data(), evecs, workspace, workspace_bytes, stream_);
    return rc == SPK_OK ? m : rc;
}

extern "C" int64_t spk_kmeans_workspace_bytes(int64_t N, int32_t d, int32_t k) {
    (void)N;
    return align_up((int64_t)k * d * 4, 256) + 256;
}

extern "C" int spk_kmeans(const float *pts, int64_t N, int32_t d, int32_t k, const float *init_centres_host, int32_t max_iter,
                          float tol, int32_t *labels, float *inertia_host, void *workspace, int64_t workspace_bytes,
                          void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(pts != nullptr && init_centres_host != nullptr && labels != nullptr, "null buffer");
    SPK_REQUIRE(N >= 1 && N < (1ll << 31) && d >= 1 && k >= 1 && k <= 64 && d <= 64, "bad shape N=%lld d=%d k=%d",
                (long long)N, d, k);
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    SPK_REQUIRE(workspace != nullptr && workspace_bytes >= spk_kmeans_workspace_bytes(N, d, k), "workspace too small");
    char *ws = static_cast<char *>(workspace);
    const int64_t cb = align_up((int64_t)k * d * 4, 256);
    float *C = reinterpret_cast<float *>(ws);
    float *scal = reinterpret_cast<float *>(ws + cb);        // [0] inertia, [1] centre shift^2, [2] (int) changed labels
    SPK_CUDA_OK(cudaMemcpyAsync(C, init_centres_host, (size_t)k * d * 4, cudaMemcpyHostToDevice, s));
    SPK_CUDA_OK(cudaMemsetAsync(labels, 0xFF, (size_t)N * 4, s));
    const int threads = 256, blocks = (int)((N + threads - 1) / threads);
    const size_t sh = (size_t)k * d * sizeof(float);
    int it = 0;
    float host_scal[3];
    bool converged = false;
    while (it < max_iter && !converged) {
        SPK_CUDA_OK(cudaMemsetAsync(scal, 0, 16, s));
        kmeans_assign_kernel<<<blocks, threads, sh, s>>>(pts, (int)N, d, k, C, labels, scal, reinterpret_cast<int *>(scal + 2));
        kmeans_update_kernel<<<k, 256, 0, s>>>(pts, (int)N, d, labels, C, scal + 1);
        count_launch(2);
        SPK_CUDA_OK(cudaMemcpyAsync(host_scal, scal, 12, cudaMemcpyDeviceToHost, s));
        SPK_CUDA_OK(cudaStreamSynchronize(s));
        int changed;
        memcpy(&changed, &host_scal[2], 4);
        ++it;
        converged = (changed == 0) || (host_scal[1] <= tol);
    }
    // final labelling against the last centres (sklearn re-labels after the last centre update)
    SPK_CUDA_OK(cudaMemsetAsync(scal, 0, 16, s));
    kmeans_assign_kernel<<<blocks, threads, sh, s>>>(pts, (int)N, d, k, C, labels, scal, reinterpret_cast<int *>(scal + 2));
    rc = check_launch("kmeans_assign_kernel");
    if (rc != SPK_OK) return rc;
    SPK_CUDA_OK(cudaMemcpyAsync(host_scal, scal, 12, cudaMemcpyDeviceToHost, s));
    SPK_CUDA_OK(cudaStreamSynchronize(s));
    if (inertia_host != nullptr) *inertia_host = host_scal[0];
    return it;
}

extern "C" int spk_cosine_pairs(const float *E, int64_t N, int64_t D, const int32_t *a, const int32_t *b, int64_t n_pairs,
                                float *out, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(n_pairs >= 0 && N >= 1 && D >= 1, "bad shape");
    if (n_pairs == 0) return SPK_OK;
    SPK_REQUIRE(E != nullptr && a != nullptr && b != nullptr && out != nullptr, "null buffer");
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const long long blocks = (n_pairs + 7) / 8;
    SPK_REQUIRE(blocks < (1ll << 31), "too many pairs for one call");
    cosine_pairs_kernel<<<(unsigned)blocks, 256, 0, s>>>(E, N, (int)D, a, b, n_pairs, out);
    return check_launch("cosine_pairs_kernel");
}

extern "C" int64_t spk_ahc_workspace_bytes(int64_t N, int64_t D) {
    if (N <= 0 || D <= 0) return 0;
    return ahc_plan(N, D).total;
}

// Average-linkage agglomerative clustering on -cos, cut at -cos_thr (AHCluster, cluster.py:139-156).
// labels: device int32 [N], numbered by the smallest member index of each cluster; returns the number of clusters.
extern "C" int spk_ahc(const float *X, int64_t N, int64_t D, float cos_thr, int32_t *labels, void *workspace,
                       int64_t workspace_bytes, void *stream_) {
    cudaStream_t s = static_cast<cudaStream_t>(stream_);
    SPK_REQUIRE(X != nullptr && labels != nullptr, "null buffer");
    SPK_REQUIRE(N >= 1 && D >= 1 && N < (1 << 24), "bad shape N=%lld D=%lld", (long long)N, (long long)D);
    int rc = require_device();
    if (rc != SPK_OK) return rc;
    const AhcPlan p = ahc_plan(N, D);
    if (workspace_bytes < p.total || workspace == nullptr) {
        set_error("workspace too small: need %lld bytes", (long long)p.total);
        return SPK_ERR_WORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    rc = cosine_matrix(X, N, D, p.aff, ws, s);
    if (rc != SPK_OK) return rc;
    float *S = reinterpret_cast<float *>(ws + p.aff.off_s);
    int *nn = reinterpret_cast<int *>(ws + p.off_nn), *size = reinterpret_cast<int *>(ws + p.off_size);
    int *parent = reinterpret_cast<int *>(ws + p.off_parent), *kdev = reinterpret_cast<int *>(ws + p.off_k);
    float *nnd = reinterpret_cast<float *>(ws + p.off_nnd);
    ahc_init_kernel<<<(unsigned)N, 256, 0, s>>>(S, (int)N, p.aff.Np, nn, nnd);
    rc = check_launch("ahc_init_kernel");
    if (rc != SPK_OK) return rc;
    ahc_merge_kernel<<<1, kAhcThreads, 0, s>>>(S, (int)N, p.aff.Np, nn, nnd, size, parent, -cos_thr, labels, kdev);
    rc = check_launch("ahc_merge_kernel");
    if (rc != SPK_OK) return rc;
    int k = 0;
    SPK_CUDA_OK(cudaMemcpyAsync(&k, kdev, sizeof(int), cudaMemcpyDeviceToHost, s));
    SPK_CUDA_OK(cudaStreamSynchronize(s));
    return k;
}
