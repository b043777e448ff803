// Micro-benchmark (debug aid): HBM -> shared memory streaming rate of TMA box loads for the access patterns of the
// bottleneck GEMM: [M rows][ld] bf16 matrix, a CTA walks tiles of 128 rows and K/64 boxes of {64 ch, 128 rows}.
// Variants change the number of boxes in flight and how many boxes of the same rows are issued back to back.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// slots: ring of `slots` buffers of 16 KB; `group` boxes (consecutive kc of the same rows) are issued per ring step
__global__ void __launch_bounds__(64, 1) stream(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, int with_b, int n_tiles, int nk, int slots, int group) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[64];
    const uint32_t b0 = smem_u32(bars);
    const int steps_per_ring = slots / group;
    if (threadIdx.x == 0) {
        for (int i = 0; i < steps_per_ring; ++i) { mbar_init(b0 + 8 * i, 1); mbar_init(b0 + 8 * (32 + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t s0 = smem_u32(smem);
    if (threadIdx.x == 0) {                  // producer
        int st = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kc = 0; kc < nk; kc += group) {
                mbar_wait(b0 + 8 * (32 + st), ph ^ 1u);
                const int n = min(group, nk - kc);
                mbar_expect(b0 + 8 * st, (uint32_t)n * 16384u * (with_b ? 2u : 1u));
                for (int e = 0; e < n; ++e) {
                    tma2d(s0 + (uint32_t)(st * group + e) * 16384u * (with_b ? 2u : 1u), &amap, (kc + e) * 64, tile * 128, b0 + 8 * st);
                    if (with_b) tma2d(s0 + (uint32_t)(st * group + e) * 32768u + 16384u, &bmap, (kc + e) * 64, 0, b0 + 8 * st);
                }
                if (++st == steps_per_ring) { st = 0; ph ^= 1u; }
            }
    } else if (threadIdx.x == 32) {          // consumer: release immediately
        int st = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kc = 0; kc < nk; kc += group) {
                mbar_wait(b0 + 8 * st, ph);
                mbar_arrive(b0 + 8 * (32 + st));
                if (++st == steps_per_ring) { st = 0; ph ^= 1u; }
            }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn enc = reinterpret_cast<EncodeFn>(p);
    const long long M = 2048LL * 74;
    const int ld = 1024;
    __nv_bfloat16 *A;
    cudaMalloc(&A, M * ld * 2);
    cudaMemset(A, 0, M * ld * 2);
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    __nv_bfloat16 *W;
    cudaMalloc(&W, 128 * 1024 * 2);
    cudaMemset(W, 0, 128 * 1024 * 2);
    const int promo = 1;
    for (int K : {512, 1024}) {
        CUtensorMap map;
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
        const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
        const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int n_tiles = (int)(M / 128), nk = K / 64;
        CUtensorMap bmap;
        const cuuint64_t bdims[2] = {(cuuint64_t)K, 128};
        const cuuint64_t bstrides[1] = {(cuuint64_t)K * 2};
        r = enc(&bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, bdims, bstrides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
        const int cfgs[][3] = {{6, 1, 0}, {6, 1, 1}, {12, 1, 0}, {6, 2, 1}};
        for (auto &c : cfgs) {
            const int slots = c[0], group = c[1], with_b = c[2];
            const int smem = slots * 16384 * (with_b ? 2 : 1);
            stream<<<148, 64, smem>>>(map, bmap, with_b, n_tiles, nk, slots, group);
            cudaEventRecord(e0);
            for (int rep = 0; rep < 3; ++rep) stream<<<148, 64, smem>>>(map, bmap, with_b, n_tiles, nk, slots, group);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            printf("K=%4d slots %2d group %d %s: %7.1f us/pass  %6.0f GB/s of A %s\n", K, slots, group, with_b ? "A+B" : "A  ",
                   ms / 3 * 1e3, (double)M * K * 2 / (ms / 3 * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
