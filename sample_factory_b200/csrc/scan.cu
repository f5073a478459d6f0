// Time-axis recurrences of the learner as warp scans (one warp per trajectory, lane = time step):
//   GAE  (algo/utils/rl_utils.py:51-94, learner.py:969-1003)   and   V-trace (learner.py:602-640).
// Both recurrences have the form  y_t = d_t + k_t * y_{t+1}, i.e. a suffix scan over the composition of affine maps
// (k, d): (k1,d1) o (k2,d2) = (k1*k2, d1 + k1*d2), done with 5 shuffle steps per 32 time steps. Row loads are one
// coalesced 128 B line per warp for T = 32 (HBM-bound: ~22 B/sample).
#include "common.cuh"

namespace sfb {

__device__ __forceinline__ void affine_suffix_scan(float& k, float& d, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ko = __shfl_down_sync(0xffffffffu, k, o);
        const float dd = __shfl_down_sync(0xffffffffu, d, o);
        if (lane + o < 32) {
            d = fmaf(k, dd, d);
            k = k * ko;
        }
    }
}

__global__ void __launch_bounds__(256) gae_returns_kernel(float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                          const uint8_t* __restrict__ time_outs,
                                                          const float* __restrict__ values,
                                                          const uint8_t* __restrict__ valids, int64_t n_traj, int T,
                                                          float gamma, float lam, int value_bootstrap,
                                                          const double* __restrict__ ret_mean,
                                                          const double* __restrict__ ret_var, float eps, float clip,
                                                          float* __restrict__ adv, float* __restrict__ returns) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool denorm = ret_mean != nullptr;
    float mu = 0.f, sigma = 1.f;
    if (denorm) {
        mu = (float)ret_mean[0];
        sigma = __fsqrt_rn(__fadd_rn((float)ret_var[0], eps));
    }
    const float gl = gamma * lam;
    const int nchunks = (T + 31) >> 5;
    for (int64_t row = warp; row < n_traj; row += nwarps) {
        const float* v = values + row * (T + 1);
        const uint8_t* vl = valids + row * (T + 1);
        float carry = 0.f;   // A_{t+1} entering the current chunk (x_last = 0, rl_utils.py:58-59)
        for (int c = nchunks - 1; c >= 0; --c) {
            const int t = (c << 5) + lane;
            const bool in = t < T;
            float k = 1.f, d = 0.f, dv_t = 0.f, valid_t = 0.f;
            if (in) {
                const int64_t j = row * T + t;
                float r = rewards[j];
                const float done = dones[j] ? 1.f : 0.f;
                float v_t = v[t], v_n = v[t + 1];
                if (denorm) {   // running_mean_std.py:107-108 (denormalize): clamp, *sigma, +mu
                    v_t = __fadd_rn(__fmul_rn(clampf(v_t, -clip, clip), sigma), mu);
                    v_n = __fadd_rn(__fmul_rn(clampf(v_n, -clip, clip), sigma), mu);
                }
                dv_t = v_t;
                valid_t = vl[t] ? 1.f : 0.f;
                const float valid_n = vl[t + 1] ? 1.f : 0.f;
                if (value_bootstrap) {   // learner.py:990
                    const float to = time_outs[j] ? 1.f : 0.f;
                    r = __fadd_rn(r, __fmul_rn(__fmul_rn(__fmul_rn(gamma, v_t), to), done));
                    rewards[j] = r;
                }
                // rl_utils.py:88: deltas = (r - v_t)*valid_t + (1-done)*(gamma*v_{t+1}*valid_{t+1})
                d = __fadd_rn(__fmul_rn(__fsub_rn(r, v_t), valid_t),
                              __fmul_rn(1.f - done, __fmul_rn(__fmul_rn(gamma, v_n), valid_n)));
                // rl_utils.py:68-69: cumulative = x + (discount*valid + (1-valid)) * cumulative * (1-done)
                k = (gl * valid_t + (1.f - valid_t)) * (1.f - done);
            }
            affine_suffix_scan(k, d, lane);
            const float a = fmaf(k, carry, d);
            carry = __shfl_sync(0xffffffffu, a, 0);
            if (in) {
                const int64_t j = row * T + t;
                adv[j] = a;
                returns[j] = __fadd_rn(a, __fmul_rn(valid_t, dv_t));   // learner.py:1003
            }
        }
    }
}

__global__ void __launch_bounds__(256) vtrace_kernel(const float* __restrict__ ratio, const float* __restrict__ values,
                                                     const float* __restrict__ rewards, const uint8_t* __restrict__ dones,
                                                     int64_t n, int R, float gamma, float rho_hat, float c_hat,
                                                     float* __restrict__ vs, float* __restrict__ adv) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nchunks = (R + 31) >> 5;
    for (int64_t row = warp; row < n; row += nwarps) {
        const int64_t base = row * R;
        // learner.py:617-619: next_values = (v[R-1] - r[R-1]) / gamma  -- the stand-in for v[R]
        const float nv_last = (values[base + R - 1] - rewards[base + R - 1]) / gamma;
        float carry = 0.f;   // u_{t+1} = vs_{t+1} - v_{t+1}; zero past the end (next_vs == next_values, :619)
        for (int c = nchunks - 1; c >= 0; --c) {
            const int t = (c << 5) + lane;
            const bool in = t < R;
            float k = 1.f, d = 0.f, rho = 0.f, r = 0.f, v_t = 0.f, nv = 0.f, ndg = 0.f;
            if (in) {
                const int64_t j = base + t;
                const float rt = ratio[j];
                rho = fminf(rho_hat, rt);                      // :613
                const float cc = fminf(c_hat, rt);             // :614
                r = rewards[j];
                v_t = values[j];
                nv = (t == R - 1) ? nv_last : values[j + 1];   // next_values of step t (:637)
                ndg = (1.f - (dones[j] ? 1.f : 0.f)) * gamma;  // :624-625
                d = rho * (r + ndg * nv - v_t);                // delta_s :631
                k = ndg * cc;                                  // :633 coefficient of (next_vs - next_values)
            }
            affine_suffix_scan(k, d, lane);
            const float u = fmaf(k, carry, d);                 // u_t = vs_t - v_t
            float u_next = __shfl_down_sync(0xffffffffu, u, 1);
            if (lane == 31) u_next = carry;
            if (in && t == R - 1) u_next = 0.f;
            carry = __shfl_sync(0xffffffffu, u, 0);
            if (in) {
                const int64_t j = base + t;
                vs[j] = v_t + u;                                          // :633-634
                adv[j] = rho * (r + ndg * (nv + u_next) - v_t);           // :632 with next_vs = nv + u_{t+1}
            }
        }
    }
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb200_gae_returns(float* rewards, const uint8_t* dones, const uint8_t* time_outs, const float* values,
                       const uint8_t* valids, int64_t n_traj, int T, float gamma, float lam, int value_bootstrap,
                       const double* ret_mean, const double* ret_var, float eps, float clip, float* adv, float* returns,
                       void* stream) {
    SFB_CHECK_ARG(rewards && dones && values && valids && adv && returns && n_traj >= 0 && T > 0,
                  "gae_returns: bad arguments");
    SFB_CHECK_ARG(!value_bootstrap || time_outs, "gae_returns: value_bootstrap needs time_outs");
    SFB_CHECK_ARG((ret_mean == nullptr) == (ret_var == nullptr), "gae_returns: ret_mean/ret_var mismatch");
    if (n_traj == 0) return 0;
    int64_t blocks = ceil_div(n_traj, 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    gae_returns_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rewards, dones, time_outs, values, valids,
                                                                           n_traj, T, gamma, lam, value_bootstrap,
                                                                           ret_mean, ret_var, eps, clip, adv, returns);
    SFB_LAUNCH_OK();
    return 0;
}

int sfb200_vtrace(const float* ratio, const float* values, const float* rewards, const uint8_t* dones, int64_t n, int R,
                  float gamma, float rho_hat, float c_hat, float* vs, float* adv, void* stream) {
    SFB_CHECK_ARG(ratio && values && rewards && dones && vs && adv && n >= 0 && R > 0, "vtrace: bad arguments");
    if (n == 0) return 0;
    int64_t blocks = ceil_div(n, 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    vtrace_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(ratio, values, rewards, dones, n, R, gamma,
                                                                      rho_hat, c_hat, vs, adv);
    SFB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
