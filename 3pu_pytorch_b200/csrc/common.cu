// Status / error text / device properties shared by every entry point of libpu3_b200.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

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

// ---- optional per-kernel-family timing inside composite entry points (level engine) -------------------------
// bench.py times kernels live with CUDA events; the engine launches ~35 kernels per call, so it records its own
// event pairs (only while enabled: two cudaEventRecord per sub-launch) and hands the sums back by tag.
struct ProfRec { int tag; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;

int prof_begin(int tag, cudaStream_t s) {
    if (!g_prof_on) return -1;
    ProfRec r; r.tag = tag;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
    cudaEventRecord(r.e0, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    return (int)g_prof.size() - 1;
}
void prof_end(int id, cudaStream_t s) {
    if (id < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (id < (int)g_prof.size()) cudaEventRecord(g_prof[id].e1, s);
}

}  // namespace pu3

extern "C" void pu3_prof_enable(int on) { pu3::g_prof_on = on != 0; }
// Adds the elapsed milliseconds / launch counts of everything recorded since the last call to ms[tag], calls[tag]
// (tags >= ntags are dropped); synchronises on the recorded events.  Returns the number of records consumed.
extern "C" int pu3_prof_collect(float *ms, int *calls, int ntags) {
    std::lock_guard<std::mutex> lk(pu3::g_prof_mu);
    int n = 0;
    for (auto &r : pu3::g_prof) {
        float t = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess &&
            r.tag >= 0 && r.tag < ntags) { ms[r.tag] += t; calls[r.tag] += 1; }
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
        ++n;
    }
    pu3::g_prof.clear();
    (void)cudaGetLastError();
    return n;
}

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
