// Micro-benchmark (debug aid): issue cost of tcgen05.mma for different issue-loop styles (N=32, no-swizzle planes).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

constexpr int N = 32;
constexpr uint32_t kPlane = 265 * 16;
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);

template <int STYLE>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long *out) {
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
    const uint32_t sa = smem_u32(smem), sb = sa + 160 * 1024;
    const uint32_t hi = ((128u >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t lo_a0 = ((sa & 0x3FFFF) >> 4) | ((kPlane >> 4) << 16);
    const uint32_t lo_b0 = ((sb & 0x3FFFF) >> 4) | ((512u >> 4) << 16);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        t0 = clock64();
        if (STYLE == 0) {                      // lane 0 branch, runtime everything (like the kernels today)
            if (threadIdx.x == 0)
                for (int i = 0; i < iters; ++i) {
                    const int k = i & 7;
                    mma(tmem, desc64(lo_a0 + ((k * 2 * kPlane) >> 4), hi), desc64(lo_b0 + ((k * 1024) >> 4), hi), kIdesc, i ? 1u : 0u);
                }
        } else if (STYLE == 1) {               // lane 0 branch, 8x unrolled with constant steps
            if (threadIdx.x == 0)
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_acc(tmem, desc64(lo_a0 + ((k * 2 * kPlane) >> 4), hi), desc64(lo_b0 + ((k * 1024) >> 4), hi), kIdesc);
                }
        } else if (STYLE == 2) {               // elect.sync once, 8x unrolled
            if (elect_one())
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma_acc(tmem, desc64(lo_a0 + ((k * 2 * kPlane) >> 4), hi), desc64(lo_b0 + ((k * 1024) >> 4), hi), kIdesc);
                }
        } else if (STYLE == 3) {               // whole warp runs the loop, each MMA predicated by elect.sync
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t da = desc64(lo_a0 + ((k * 2 * kPlane) >> 4), hi), db = desc64(lo_b0 + ((k * 1024) >> 4), hi);
                    asm volatile("{\n\t.reg .pred P, q;\n\telect.sync _|P, 0xffffffff;\n\tsetp.ne.b32 q, 1, 0;\n\t"
                                 "@P tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, q;\n\t}"
                                 ::"r"(tmem), "l"(da), "l"(db), "r"(kIdesc) : "memory");
                }
            }
        } else if (STYLE == 4) {               // same descriptors every time (no address math at all), lane 0
            if (threadIdx.x == 0) {
                const uint64_t da = desc64(lo_a0, hi), db = desc64(lo_b0, hi);
                for (int i = 0; i < iters; i += 8) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) mma_acc(tmem, da, db, kIdesc);
                }
            }
        }
        __syncwarp();
        t1 = clock64();
        if (threadIdx.x == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b) : "memory");
    }
    if (threadIdx.x == 0) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int STYLE>
void run(const char *name, long long *out) {
    const int iters = 2048;
    cudaFuncSetAttribute(bench<STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    bench<STYLE><<<148, 128, 200 * 1024>>>(iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-48s issue %6.1f cyc/MMA, complete %6.1f cyc/MMA %s\n", name, (double)h[0] / iters, (double)h[1] / iters,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long *out;
    cudaMalloc(&out, 16);
    run<0>("lane0 branch, runtime k", out);
    run<1>("lane0 branch, unrolled x8 constant steps", out);
    run<2>("elect.sync branch, unrolled x8", out);
    run<3>("warp-wide loop, per-MMA elect predicate", out);
    run<4>("lane0, loop-invariant descriptors", out);
    return 0;
}
