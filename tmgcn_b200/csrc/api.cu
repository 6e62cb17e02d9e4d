// Error plumbing, device queries and the trivial ABI entry points.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tmgcn {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int after_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

static int g_sms[64];
static size_t g_l2[64];
static bool g_have[64];

static void query() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g_have[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        g_sms[dev] = v > 0 ? v : 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev);
        g_l2[dev] = v > 0 ? (size_t)v : (size_t)126 << 20;
        g_have[dev] = true;
    }
}
int sm_count() {
    query();
    int dev = 0;
    cudaGetDevice(&dev);
    return g_sms[dev < 64 && dev >= 0 ? dev : 0];
}
size_t l2_bytes() {
    query();
    int dev = 0;
    cudaGetDevice(&dev);
    return g_l2[dev < 64 && dev >= 0 ? dev : 0];
}

}  // namespace tmgcn

extern "C" {
const char *tmgcn_last_error(void) { return tmgcn::g_err; }
int tmgcn_abi_version(void) { return TMGCN_ABI_VERSION; }
int64_t tmgcn_launch_count(void) { return tmgcn::g_launches.load(); }
}
