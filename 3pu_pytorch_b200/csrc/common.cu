// Status / error text / device properties shared by every entry point of libpu3_b200.
#include <stdarg.h>
#include <string.h>

#include "pu3_common.cuh"

namespace pu3 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

const DeviceInfo &device_info() {
    static DeviceInfo info[64];
    static bool have[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!have[dev]) {
        DeviceInfo d{148, 232448, 100};  // B200 values, used when no device can be queried
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) d.sm_count = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0) d.smem_optin = v;
        int maj = 0, mnr = 0;
        if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev) == cudaSuccess && maj > 0)
            d.cc = maj * 10 + mnr;
        (void)cudaGetLastError();
        info[dev] = d;
        have[dev] = true;
    }
    return info[dev];
}

}  // namespace pu3

extern "C" const char *pu3_last_error(void) { return pu3::g_err; }
extern "C" int pu3_version(void) { return 1; }
extern "C" int pu3_device_info(int *sm_count, int *smem_optin_bytes, int *cc) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        pu3::set_error("no CUDA device visible: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return e == cudaSuccess ? PU3_E_UNSUPPORTED : (int)e;
    }
    const pu3::DeviceInfo &d = pu3::device_info();
    if (sm_count) *sm_count = d.sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = d.smem_optin;
    if (cc) *cc = d.cc;
    return PU3_OK;
}
