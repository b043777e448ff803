// Micro-benchmark (debug aid): cycles per tcgen05.mma for operand layouts / shapes used by the conv kernels.
// One CTA per SM; lane 0 of warp 0 issues `iters` MMAs back to back and waits for the commit.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }

struct Variant {
    const char *name;
    int N, layout;           // layout: 0 none, 1 = 128B swizzle, 2 = 64B, 3 = 32B  (descriptor bits 61-63: 0,2,4,6)
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_kstep, b_kstep;      // bytes added to the start address per MMA
    int ksteps;                     // distinct K steps cycled through
    int a_mn_major;
    int n_acc;                      // accumulators cycled through
    uint32_t a_tile_step;           // bytes between A tiles when cycling accumulators
};

__global__ void __launch_bounds__(128, 1) bench(Variant v, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem), sb = sa + 160 * 1024;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)v.a_mn_major << 15) | ((uint32_t)(v.N >> 3) << 17) | (8u << 24);
        const uint32_t lay = v.layout == 1 ? 2u : v.layout == 2 ? 4u : v.layout == 3 ? 6u : 0u;
        const uint32_t hi_a = ((v.a_sbo >> 4) & 0x3FFF) | (1u << 14) | (lay << 29);
        const uint32_t hi_b = ((v.b_sbo >> 4) & 0x3FFF) | (1u << 14) | (lay << 29);
        const uint32_t lo_a0 = ((sa & 0x3FFFF) >> 4) | ((v.a_lbo >> 4) << 16);
        const uint32_t lo_b0 = ((sb & 0x3FFFF) >> 4) | ((v.b_lbo >> 4) << 16);
        long long t0 = clock64();
        int k = 0, acc = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t lo_a = lo_a0 + ((k * v.a_kstep + acc * v.a_tile_step) >> 4), lo_b = lo_b0 + ((k * v.b_kstep) >> 4);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + acc * v.N), "l"(desc64(lo_a, hi_a)), "l"(desc64(lo_b, hi_b)), "r"(idesc), "r"(i >= v.n_acc ? 1u : 0u) : "memory");
            if (++acc == v.n_acc) { acc = 0; if (++k == v.ksteps) k = 0; }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    const uint32_t plane = 265 * 16;
    Variant vs[] = {
        // name,                         N  lay  a_lbo  a_sbo b_lbo b_sbo  a_k      b_k  ks mn acc tile
        {"nosw planes N=32 (slab)",      32, 0, plane,  128,  512,  128, 2 * plane, 1024, 8, 0, 1, 0},
        {"nosw planes N=32 2acc",        32, 0, plane,  128,  512,  128, 2 * plane, 1024, 8, 0, 2, 2048},
        {"nosw planes N=64",             64, 0, plane,  128, 1024,  128, 2 * plane, 2048, 8, 0, 1, 0},
        {"nosw planes N=128",           128, 0, plane,  128, 2048,  128, 2 * plane, 4096, 4, 0, 1, 0},
        {"nosw packed (lbo=2048) N=32",  32, 0, 2048,   128,  512,  128, 4096,      1024, 8, 0, 1, 0},
        {"nosw same-K N=32",             32, 0, plane,  128,  512,  128, 0,         0,    1, 0, 1, 0},
        {"sw128 N=32",                   32, 1, 16,    1024,   16, 1024, 32,        32,   4, 0, 1, 0},
        {"sw128 N=32 2acc",              32, 1, 16,    1024,   16, 1024, 32,        32,   4, 0, 2, 16384},
        {"sw128 N=64",                   64, 1, 16,    1024,   16, 1024, 32,        32,   4, 0, 1, 0},
        {"sw128 N=128",                 128, 1, 16,    1024,   16, 1024, 32,        32,   4, 0, 1, 0},
        {"sw128 N=256",                 256, 1, 16,    1024,   16, 1024, 32,        32,   4, 0, 1, 0},
        {"sw64 N=32",                    32, 2, 16,     512,   16,  512, 32,        32,   2, 0, 1, 0},
        {"sw32 N=32",                    32, 3, 16,     256,   16,  256, 0,         0,    1, 0, 1, 0},
        {"nosw MN-major A N=16 (sums)",  16, 0, 128,  plane,  256,  128, 256,       512,  8, 1, 1, 0},
    };
    long long *out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2048;
    for (const Variant &v : vs) {
        for (int grid : {1, 148}) {
            bench<<<grid, 128, 200 * 1024>>>(v, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2] = {0, 0};
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("%-32s grid %3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (floor %d)  %s\n", v.name, grid, (double)h[0] / iters,
                   (double)h[1] / iters, 128 * v.N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
