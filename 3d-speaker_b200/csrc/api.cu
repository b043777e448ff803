// Library-level entry points: error string, ABI version, device check, launch counter.
#include <mutex>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace spk {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

static int check_dev(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); libb200spk has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return SPK_ERR_NO_DEVICE;
    }
    if (dev < 0 || dev >= n) {
        set_error("device %d out of range (%d devices)", dev, n);
        return SPK_ERR_INVALID;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
        return SPK_ERR_CUDA;
    }
    if (major != 10) {
        set_error("device %d is sm_%dx; libb200spk is built for sm_100a only", dev, major);
        return SPK_ERR_NO_DEVICE;
    }
    return SPK_OK;
}

int require_device() {
    // cached per device id; cudaGetDevice is cheap
    static std::mutex mu;
    static int ok_mask = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); libb200spk has no CPU fallback", cudaGetErrorString(e));
        return SPK_ERR_NO_DEVICE;
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 31 && (ok_mask >> dev) & 1) return SPK_OK;
    }
    int rc = check_dev(dev);
    if (rc == SPK_OK && dev < 31) {
        std::lock_guard<std::mutex> lk(mu);
        ok_mask |= 1 << dev;
    }
    return rc;
}

bool nvtx_enabled() {
    static const bool on = [] { const char *e = getenv("SPK_NVTX"); return e && e[0] == '1'; }();
    return on;
}
void nvtx_push(const char *name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

int sm_count() {
    static int cached[32] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 32 && cached[dev]) return cached[dev];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 32) cached[dev] = n;
    return n;
}

}  // namespace spk

extern "C" int spk_abi_version(void) { return SPK_ABI_VERSION; }
extern "C" const char *spk_last_error(void) { return spk::t_err; }
extern "C" int spk_device_check(int dev) { return spk::check_dev(dev); }
extern "C" int64_t spk_launch_count(void) { return spk::g_launches.load(); }
extern "C" void spk_add_launches(int64_t n) { spk::g_launches.fetch_add(n, std::memory_order_relaxed); }
