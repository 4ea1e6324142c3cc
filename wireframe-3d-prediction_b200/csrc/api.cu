// Library-wide plumbing: thread-local error string, version, device info.
#include "wf_common.cuh"

#include <mutex>

namespace wf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace wf

extern "C" int wf_version(void) { return 100; }
extern "C" const char* wf_last_error(void) { return wf::g_err; }
extern "C" int wf_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    WF_CUDA(cudaGetDevice(&dev));
    if (sm) WF_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
    if (major) WF_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
    if (minor) WF_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
    return WF_OK;
}
