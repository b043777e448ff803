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

#include "ops.cuh"

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

// ---- Lanczos building blocks (fp32 storage, fp32 accumulate with tree reductions)
// w = sigma*v - L v      (one warp per row)
__global__ void __launch_bounds__(256)
shifted_matvec_kernel(const float *__restrict__ L, int N, int ld, float sigma, const float *__restrict__ v,
                      float *__restrict__ w) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const float *lr = L + (size_t)row * ld;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s = fmaf(lr[j], v[j], s);
    s = warp_sum(s);
    if (lane == 0) w[row] = sigma * v[row] - s;
}
// h[j] = dot(V[j,:], w) for j < m     (one CTA per j)
__global__ void __launch_bounds__(256)
dots_kernel(const float *__restrict__ V, int N, int m, const float *__restrict__ w, float *__restrict__ h) {
    __shared__ float sh[32];
    const int j = blockIdx.x;
    const float *vj = V + (size_t)j * N;
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s = fmaf(vj[i], w[i], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) h[j] = s;
}
// w -= sum_j h[j] V[j,:] in two deterministic stages: gridDim.y slices of the basis produce partial
// sums (so the serial depth per thread is m / gridDim.y, not m), a second kernel folds them in order
__global__ void __launch_bounds__(256)
axpys_partial_kernel(const float *__restrict__ V, int N, int m, const float *__restrict__ h, float *__restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int per = (m + gridDim.y - 1) / gridDim.y;
    const int j0 = blockIdx.y * per, j1 = min(m, j0 + per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int j = j0;
    for (; j + 3 < j1; j += 4) {
        s0 = fmaf(h[j], V[(size_t)j * N + i], s0);
        s1 = fmaf(h[j + 1], V[(size_t)(j + 1) * N + i], s1);
        s2 = fmaf(h[j + 2], V[(size_t)(j + 2) * N + i], s2);
        s3 = fmaf(h[j + 3], V[(size_t)(j + 3) * N + i], s3);
    }
    for (; j < j1; ++j) s0 = fmaf(h[j], V[(size_t)j * N + i], s0);
    part[(size_t)blockIdx.y * N + i] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256)
axpys_finish_kernel(const float *__restrict__ part, int N, int slices, float *__restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float s = 0.f;
    for (int y = 0; y < slices; ++y) s += part[(size_t)y * N + i];
    w[i] -= s;
}
// beta = ||w||; V[m,:] = w / beta; record alpha (sum of the two projections on v_{m-1}) and beta
__global__ void __launch_bounds__(256)
normalize_next_kernel(const float *__restrict__ w, int N, float *__restrict__ vnext, const float *__restrict__ h1,
                      const float *__restrict__ h2, int jlast, float *__restrict__ alpha, float *__restrict__ beta, int step) {
    __shared__ float sh[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s = fmaf(w[i], w[i], s);
    s = block_sum(s, sh);
    const float b = sqrtf(s);
    const float inv = b > 0.f ? 1.f / b : 0.f;
    if (vnext != nullptr)
        for (int i = threadIdx.x; i < N; i += blockDim.x) vnext[i] = w[i] * inv;
    if (threadIdx.x == 0) {
        alpha[step] = h1[jlast] + h2[jlast];
        beta[step] = b;
    }
}
__global__ void init_vector_kernel(float *__restrict__ v, int N, unsigned seed) {
    __shared__ float sh[32];
    // deterministic pseudo-random start vector (hash), then normalised by the caller kernel
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u ^ seed;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        v[i] = ((x >> 8) * (1.0f / 16777216.0f)) - 0.5f;
    }
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s = fmaf(v[i], v[i], s);
    s = block_sum(s, sh);
    const float inv = rsqrtf(s);
    for (int i = threadIdx.x; i < N; i += blockDim.x) v[i] *= inv;
}
// evecs[i, c] = sum_j V[j, i] * Sm[j, c]     (Ritz vectors; Sm is m x k row-major on device)
__global__ void __launch_bounds__(256)
ritz_kernel(const float *__restrict__ V, int N, int m, const float *__restrict__ Sm, int k, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    for (int j = 0; j < m; ++j) {
        const float v = V[(size_t)j * N + i];
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (c < k) acc[c] = fmaf(v, Sm[j * k + c], acc[c]);
    }
    for (int c = 0; c < k; ++c) out[(size_t)i * k + c] = acc[c];
}
__global__ void gershgorin_kernel(const float *__restrict__ L, int N, int ld, float *__restrict__ out) {
    // sigma = max_i 2*L[i][i] (unnormalised Laplacian: row sum of |offdiag| equals the diagonal)
    __shared__ float sh[32];
    float mx = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, L[(size_t)i * ld + i]);
    // block max via the sum helper on a one-hot trick is wasteful; do a plain shared reduction
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sh[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m2 = 0.f;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m2 = fmaxf(m2, sh[w]);
        out[0] = 2.f * m2 + 1e-3f;
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
constexpr int kAxpySlices = 16;
struct EigPlan { int m_max; int64_t off_V, off_w, off_part, off_h1, off_h2, off_alpha, off_beta, off_sm, off_sigma, total; };
EigPlan eig_plan(int64_t N, int k) {
    EigPlan p;
    // Krylov budget: clustered eigenvalues at the edge of the bulk (and exact multiplicities, which a
    // single-vector Lanczos only separates through round-off) can need several hundred steps
    p.m_max = (int)std::min<int64_t>(N, std::max(48 * k, 1024));
    int64_t cur = 0;
    auto take = [&](int64_t bytes) { int64_t o = cur; cur += align_up(bytes, 256); return o; };
    p.off_V = take((int64_t)(p.m_max + 1) * N * 4);
    p.off_w = take(N * 4);
    p.off_part = take((int64_t)kAxpySlices * N * 4);
    p.off_h1 = take((int64_t)(p.m_max + 1) * 4);
    p.off_h2 = take((int64_t)(p.m_max + 1) * 4);
    p.off_alpha = take((int64_t)(p.m_max + 1) * 4);
    p.off_beta = take((int64_t)(p.m_max + 1) * 4);
    p.off_sm = take((int64_t)p.m_max * 32 * 4);
    p.off_sigma = take(256);
    p.total = cur;
    return p;
}
}  // namespace

extern "C" int64_t spk_eig_workspace_bytes(int64_t N, int32_t k) {
    if (N <= 0 || k <= 0) return 0;
    return eig_plan(N, k).total;
}

namespace {
struct EigBufs { float *V, *w, *part, *h1, *h2, *alpha, *beta, *Sm, *dsigma; };
EigBufs eig_bufs(void *workspace, const EigPlan &p) {
    char *ws = static_cast<char *>(workspace);
    EigBufs b;
    b.V = reinterpret_cast<float *>(ws + p.off_V); b.w = reinterpret_cast<float *>(ws + p.off_w);
    b.part = reinterpret_cast<float *>(ws + p.off_part);
    b.h1 = reinterpret_cast<float *>(ws + p.off_h1); b.h2 = reinterpret_cast<float *>(ws + p.off_h2);
    b.alpha = reinterpret_cast<float *>(ws + p.off_alpha); b.beta = reinterpret_cast<float *>(ws + p.off_beta);
    b.Sm = reinterpret_cast<float *>(ws + p.off_sm); b.dsigma = reinterpret_cast<float *>(ws + p.off_sigma);
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
        gershgorin_kernel<<<1, 256, 0, s>>>(L, n, ld, b.dsigma);
        init_vector_kernel<<<1, 256, 0, s>>>(b.V, n, 0x9E3779B9u);
        count_launch(2);
    }
    SPK_CUDA_OK(cudaMemcpyAsync(sigma_host, b.dsigma, sizeof(float), cudaMemcpyDeviceToHost, s));
    SPK_CUDA_OK(cudaStreamSynchronize(s));
    const float sigma = *sigma_host;
    const int gx = (n + 255) / 256;
    for (int j = m_from; j < m_to; ++j) {
        float *vj = b.V + (size_t)j * n;
        shifted_matvec_kernel<<<(n + 7) / 8, 256, 0, s>>>(L, n, ld, sigma, vj, b.w);
        // full re-orthogonalisation: classical Gram-Schmidt applied twice
        const int slices = std::min(kAxpySlices, (j + 1 + 31) / 32);
        dots_kernel<<<j + 1, 256, 0, s>>>(b.V, n, j + 1, b.w, b.h1);
        axpys_partial_kernel<<<dim3(gx, slices), 256, 0, s>>>(b.V, n, j + 1, b.h1, b.part);
        axpys_finish_kernel<<<gx, 256, 0, s>>>(b.part, n, slices, b.w);
        dots_kernel<<<j + 1, 256, 0, s>>>(b.V, n, j + 1, b.w, b.h2);
        axpys_partial_kernel<<<dim3(gx, slices), 256, 0, s>>>(b.V, n, j + 1, b.h2, b.part);
        axpys_finish_kernel<<<gx, 256, 0, s>>>(b.part, n, slices, b.w);
        normalize_next_kernel<<<1, 256, 0, s>>>(b.w, n, b.V + (size_t)(j + 1) * n, b.h1, b.h2, j, b.alpha, b.beta, j);
        count_launch(8);
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        set_error("lanczos launch failed: %s", cudaGetErrorString(le));
        return SPK_ERR_CUDA;
    }
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
    ritz_kernel<<<((int)N + 255) / 256, 256, 0, s>>>(b.V, (int)N, m, b.Sm, k, evecs);
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
