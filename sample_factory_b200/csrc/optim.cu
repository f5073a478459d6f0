// Global grad-norm clip + Adam on flat parameter / gradient / moment buffers (learner.py:782-797; torch
// clip_grad_norm_ and torch.optim.Adam semantics, SURVEY App.A-12).  HBM-bound: 28 B/param (+4 B for the norm pass).
#include <math.h>

#include "common.cuh"

namespace sfb {

constexpr int kNormBlocks = 480;   // partial slots in the 4 KiB workspace (doubles)

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ part) {
    __shared__ double sm[8];
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = (double)g[i];
        s += v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < 8 ? sm[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) part[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        double lr, double bc1, float bc2_sqrt, float omb1, float beta2,
                                                        float omb2,
                                                        float eps, float max_norm, const double* __restrict__ part,
                                                        int nparts, const double* __restrict__ lr_num,
                                                        const double* __restrict__ lr_den,
                                                        float* __restrict__ grad_norm_out, float* __restrict__ p_lo) {
    __shared__ float s_coef;
    __shared__ float s_step;
    if (threadIdx.x < 32) {
        // every block reduces the (<= 480) partials itself, in the same order -> identical coefficient everywhere
        double t = 0.0;
        for (int k = threadIdx.x; k < nparts; k += 32) t += part[k];
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            const float total = (float)sqrt(t);
            float coef = 1.f;
            if (max_norm > 0.f) coef = fminf(__fdiv_rn(max_norm, total + 1e-6f), 1.0f);   // clip_grad.py
            s_coef = coef;
            double lr_eff = lr;
            if (lr_num && lr_den) lr_eff = lr * lr_num[0] / lr_den[0];                     // learner.py:788-794
            s_step = (float)(lr_eff / bc1);                                                 // adam.py step_size
            if (grad_norm_out && blockIdx.x == 0) grad_norm_out[0] = total;
        }
    }
    __syncthreads();
    const float coef = s_coef, step_size = s_step;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float mi = m[i], vi = v[i];
        mi = mi + omb1 * (gi - mi);                        // exp_avg.lerp_(grad, 1-beta1)
        vi = vi * beta2 + (omb2 * gi) * gi;                // exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2)
        const float denom = __fdiv_rn(__fsqrt_rn(vi), bc2_sqrt) + eps;
        const float pn = p[i] - step_size * __fdiv_rn(mi, denom);   // param.addcdiv_(exp_avg, denom, value=-step_size)
        p[i] = pn;
        if (p_lo) p_lo[i] = __uint_as_float(tf32_lo_bits(__float_as_uint(pn)));   // registered tf32 low half stays current
        m[i] = mi;
        v[i] = vi;
    }
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb200_clip_adam_step(float* p, float* g, float* m, float* v, int64_t n, int64_t step, double lr, double beta1,
                          double beta2, double eps, double max_grad_norm, const double* lr_scale_num,
                          const double* lr_scale_den, float* grad_norm_out, void* workspace, void* stream) {
    SFB_CHECK_ARG(p && g && m && v && workspace && n > 0 && step >= 1, "clip_adam_step: bad arguments");
    SFB_CHECK_ARG((lr_scale_num == nullptr) == (lr_scale_den == nullptr), "clip_adam_step: lr_scale num/den mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    double* part = (double*)workspace;
    int64_t nb = ceil_div(n, 256 * 4);
    if (nb > kNormBlocks) nb = kNormBlocks;
    sumsq_kernel<<<(unsigned)nb, 256, 0, st>>>(g, n, part);
    SFB_LAUNCH_OK();
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    int64_t blocks = ceil_div(n, 256 * 2);
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    clip_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, lr, bc1, (float)sqrt(bc2), (float)(1.0 - beta1),
                                                       (float)beta2, (float)(1.0 - beta2), (float)eps, (float)max_grad_norm, part, (int)nb,
                                                       lr_scale_num, lr_scale_den, grad_norm_out, tf32_lo_lookup_mut(p, n));
    SFB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
