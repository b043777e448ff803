// Shared helpers for libb200spk (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/b200spk.h"

namespace spk {

// ---- per-thread error message + global launch counter (api.cu)
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SPK_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            spk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                           __LINE__);                                                       \
            return SPK_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define SPK_REQUIRE(cond, ...)           \
    do {                                 \
        if (!(cond)) {                   \
            spk::set_error(__VA_ARGS__); \
            return SPK_ERR_INVALID;      \
        }                                \
    } while (0)

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    count_launch();
    return SPK_OK;
}

int require_device();   // SPK_OK when the current device is sm_100 (cached)
int sm_count();

// ---- dtype helpers
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == SPK_ACT_RELU) return fmaxf(v, 0.f);
    if (act == SPK_ACT_CLAMP20) return fminf(fmaxf(v, 0.f), 20.f);
    if (act == SPK_ACT_SILU) return v / (1.f + __expf(-v));
    if (act == SPK_ACT_TANH) return tanhf(v);
    return v;
}

// the same on a register array, with the (warp-uniform) activation switch taken once instead of per element
template <int N>
__device__ __forceinline__ void apply_act_vec(float (&v)[N], int act) {
    if (act == SPK_ACT_RELU) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = fmaxf(v[e], 0.f);
    } else if (act == SPK_ACT_CLAMP20) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = fminf(fmaxf(v[e], 0.f), 20.f);
    } else if (act == SPK_ACT_SILU) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = v[e] / (1.f + __expf(-v[e]));
    } else if (act == SPK_ACT_TANH) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = tanhf(v[e]);
    }
}

// ---- programmatic dependent launch: the grid may start while the previous kernel of the stream drains.  Kernels
// launched this way do their input-independent prologue (barrier init, TMEM allocation, weight staging), then
// pdl_wait() before the first read of anything an earlier kernel wrote; pdl_trigger() at their own start lets the
// next kernel do the same.  SPK_NO_PDL=1 falls back to plain stream order.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("SPK_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---- NVTX ranges (header-only nvtx3; a no-op unless a profiler is attached).  SPK_NVTX=1 switches them on: one range
// per library call and per fused op, so an nsys / ncu timeline reads as "fbank | forward(stem, conv, cam_local, ...) |
// affinity | lanczos | kmeans" instead of anonymous kernels.
bool nvtx_enabled();
void nvtx_push(const char *name);
void nvtx_pop();
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char *name) : on(nvtx_enabled()) { if (on) nvtx_push(name); }
    ~NvtxRange() { if (on) nvtx_pop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

}  // namespace spk
