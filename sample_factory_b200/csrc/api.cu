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

struct LoRange {
    const float* base;
    float* lo;
    int64_t n;
};
static LoRange g_lo[32];
static int g_lo_n = 0;

float* tf32_lo_lookup_mut(float* p, int64_t count) {
    for (int i = 0; i < g_lo_n; ++i) {
        const LoRange& r = g_lo[i];
        if (p >= r.base && p + count <= r.base + r.n) return r.lo + (p - r.base);
    }
    return nullptr;
}
const float* tf32_lo_lookup(const float* p, int64_t count) { return tf32_lo_lookup_mut(const_cast<float*>(p), count); }

bool tf32_lo_check_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SFB200_CHECK_LO");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

__global__ void tf32_split_lo_kernel(const float* __restrict__ w, float* __restrict__ lo, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        lo[i] = __uint_as_float(tf32_lo_bits(__float_as_uint(w[i])));
}

__global__ void tf32_check_lo_kernel(const float* __restrict__ w, const float* __restrict__ lo, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (__float_as_uint(lo[i]) != tf32_lo_bits(__float_as_uint(w[i]))) {
            printf("libsfb200: STALE tf32 lo buffer at element %lld (call sfb200_refresh_tf32_lo after writing weights)\n",
                   (long long)i);
            __trap();
        }
}

int tf32_lo_check(const float* w, const float* lo, int64_t count, cudaStream_t st) {
    int64_t blocks = ceil_div(count, 256);
    if (blocks > 1024) blocks = 1024;
    tf32_check_lo_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, lo, count);
    SFB_LAUNCH_OK();
    return 0;
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

int sfb200_register_tf32_lo(const float* base, float* lo, int64_t n) {
    SFB_CHECK_ARG(base && lo && n > 0, "register_tf32_lo: bad arguments");
    for (int i = 0; i < sfb::g_lo_n; ++i)
        if (sfb::g_lo[i].base == base) {
            sfb::g_lo[i] = sfb::LoRange{base, lo, n};
            return 0;
        }
    SFB_CHECK_ARG(sfb::g_lo_n < 32, "register_tf32_lo: table full");
    sfb::g_lo[sfb::g_lo_n++] = sfb::LoRange{base, lo, n};
    return 0;
}

int sfb200_unregister_tf32_lo(const float* base) {
    for (int i = 0; i < sfb::g_lo_n; ++i)
        if (sfb::g_lo[i].base == base) {
            sfb::g_lo[i] = sfb::g_lo[--sfb::g_lo_n];
            return 0;
        }
    return 0;
}

int sfb200_refresh_tf32_lo(const float* base, void* stream) {
    for (int i = 0; i < sfb::g_lo_n; ++i)
        if (sfb::g_lo[i].base == base) {
            const sfb::LoRange& r = sfb::g_lo[i];
            int64_t blocks = sfb::ceil_div(r.n, 256);
            if (blocks > 1184) blocks = 1184;
            sfb::tf32_split_lo_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(r.base, r.lo, r.n);
            SFB_LAUNCH_OK();
            return 0;
        }
    sfb::set_error("refresh_tf32_lo: buffer is not registered");
    return 1;
}

uint64_t sfb200_launch_count(void) { return __atomic_load_n(&sfb::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"
