// Experiment (debug aid): do swizzled K-major UMMA descriptors work when the start address is shifted by whole
// rows (not a multiple of the 8-row swizzle atom)?  Needed for "taps = shifted descriptors" on TMA-friendly layouts.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }

constexpr int ROWS = 160, N = 32;

// mode: 1 = SW128 (row 128 B = 64 bf16), 2 = SW64 (row 64 B = 32 bf16)
__global__ void __launch_bounds__(128, 1) test(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D, int mode, int shift, int base_off) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int row_bytes = mode == 1 ? 128 : 64, kcols = row_bytes / 2, chunks = row_bytes / 16;
    uint8_t *sA = smem, *sB = smem + 32 * 1024;
    // swizzle on absolute (1024-aligned base) addresses: 16-byte chunk index ^= (row % 8) [SW128] / ((row/2) % 4) [SW64]
    for (int i = threadIdx.x; i < ROWS * chunks; i += blockDim.x) {
        const int r = i / chunks, c = i % chunks;
        const int phys = mode == 1 ? (c ^ (r & 7)) : (c ^ ((r >> 1) & 3));
        *reinterpret_cast<uint4 *>(sA + r * row_bytes + phys * 16) = *reinterpret_cast<const uint4 *>(A + r * kcols + c * 8);
    }
    for (int i = threadIdx.x; i < N * chunks; i += blockDim.x) {
        const int r = i / chunks, c = i % chunks;
        const int phys = mode == 1 ? (c ^ (r & 7)) : (c ^ ((r >> 1) & 3));
        *reinterpret_cast<uint4 *>(sB + r * row_bytes + phys * 16) = *reinterpret_cast<const uint4 *>(B + r * kcols + c * 8);
    }
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t lay = mode == 1 ? 2u : 4u;
        const uint32_t sbo = 8 * row_bytes;
        const uint32_t hi_a = ((sbo >> 4) & 0x3FFF) | (1u << 14) | ((uint32_t)(base_off & 7) << 17) | (lay << 29);
        const uint32_t hi_b = ((sbo >> 4) & 0x3FFF) | (1u << 14) | (lay << 29);
        const uint32_t a0 = smem_u32(sA) + shift * row_bytes, b0 = smem_u32(sB);
        for (int k = 0; k < kcols / 16; ++k) {
            const uint32_t lo_a = (((a0 + k * 32) & 0x3FFFF) >> 4) | (1u << 16), lo_b = (((b0 + k * 32) & 0x3FFFF) >> 4) | (1u << 16);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(desc64(lo_a, hi_a)), "l"(desc64(lo_b, hi_b)), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t r[32];
    for (int h = 0; h < 2; ++h) {
        uint32_t *q = r + 16 * h;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]),
                       "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + 16 * h) : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 32; ++e) D[(warp * 32 + lane) * N + e] = __uint_as_float(r[e]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    for (int mode = 1; mode <= 2; ++mode) {
        const int kcols = mode == 1 ? 64 : 32;
        __nv_bfloat16 *hA = new __nv_bfloat16[ROWS * kcols], *hB = new __nv_bfloat16[N * kcols];
        float *fA = new float[ROWS * kcols], *fB = new float[N * kcols];
        srand(7);
        for (int i = 0; i < ROWS * kcols; ++i) { fA[i] = (float)(rand() % 9 - 4); hA[i] = __float2bfloat16(fA[i]); }
        for (int i = 0; i < N * kcols; ++i) { fB[i] = (float)(rand() % 7 - 3); hB[i] = __float2bfloat16(fB[i]); }
        __nv_bfloat16 *dA, *dB; float *dD;
        cudaMalloc(&dA, ROWS * kcols * 2); cudaMalloc(&dB, N * kcols * 2); cudaMalloc(&dD, 128 * N * 4);
        cudaMemcpy(dA, hA, ROWS * kcols * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB, N * kcols * 2, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        float hD[128 * N];
        for (int shift = 0; shift <= 9; ++shift)
            for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
                const int base_off = bo_mode ? (mode == 1 ? (shift & 7) : ((shift >> 1) & 3)) : 0;
                if (bo_mode && base_off == 0) continue;
                test<<<1, 128, 64 * 1024>>>(dA, dB, dD, mode, shift, base_off);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d shift %d base_off %d: %s\n", mode, shift, base_off, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
                double worst = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < kcols; ++k) ref += (double)fA[(m + shift) * kcols + k] * fB[n * kcols + k];
                        worst = fmax(worst, fabs(ref - hD[m * N + n]));
                    }
                printf("%s shift %d rows, base_offset %d: max |err| = %g %s\n", mode == 1 ? "SW128" : "SW64 ", shift, base_off, worst, worst == 0 ? "OK" : "WRONG");
            }
    }
    return 0;
}
