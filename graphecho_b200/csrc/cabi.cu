// Library-level C-ABI entry points: version, last-error text, device facts.
#include "common.cuh"
#include "../../include/graphecho_b200.h"
#include <mutex>

namespace {
thread_local char g_err[512] = "";
}

void ge_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace ge {
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
}  // namespace ge

extern "C" int ge_version(void) { return GE_ABI_VERSION; }

extern "C" const char* ge_last_error(void) { return g_err; }

extern "C" int ge_device_sm_count(void) { return ge::sm_count(); }

// Number of kernels this library has launched on behalf of the caller since load
// (bench.py reports it as gpu_launches).
namespace {
unsigned long long g_launches = 0;
}
extern "C" unsigned long long ge_launch_count(void) { return g_launches; }
extern "C" void ge_count_launches(unsigned long long n) { __atomic_add_fetch(&g_launches, n, __ATOMIC_RELAXED); }
