// api.cu -- library-wide C ABI helpers (version, errors, launch accounting, device info).
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace piml {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PIML_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return PIML_OK;
}

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;          // B200; not cached so a later call can still succeed
    }
    return cached;
}

}  // namespace piml

extern "C" int piml_version(void) { return 100; }

extern "C" const char *piml_last_error(void) { return piml::g_err; }

extern "C" int64_t piml_launch_count(void) { return piml::g_launches.load(std::memory_order_relaxed); }

extern "C" int piml_device_info(int *sm_count, int *cc) {
    int dev = 0, n = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -piml::fail(PIML_ERR_CUDA, "no CUDA device");
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (sm_count) *sm_count = n;
    if (cc) *cc = major * 10 + minor;
    return 0;
}
