// Library-level entry points of libsfb200 (see include/sfb200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sfb {
static thread_local char g_err[512] = "";
static thread_local int g_sm_count = 0;

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SFB200_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            int n = 0;
            if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_sm_count = n;
        }
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}
}  // namespace sfb

extern "C" {

int sfb200_abi_version(void) { return SFB200_ABI_VERSION; }

const char* sfb200_last_error(void) { return sfb::g_err; }

int sfb200_set_device(int device) {
    SFB_CUDA_OK(cudaSetDevice(device));
    sfb::g_sm_count = 0;
    return 0;
}

int sfb200_sm_count(void) { return sfb::sm_count(); }

uint64_t sfb200_launch_count(void) { return __atomic_load_n(&sfb::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"
