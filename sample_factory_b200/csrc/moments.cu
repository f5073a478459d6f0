// Running-moment kernels (algo/utils/running_mean_std.py:49-77): column-wise batch mean / unbiased variance with fp64
// accumulation (HBM-bound: one pass over x), and the fp64 Welford merge into the running buffers.
#include "common.cuh"

namespace sfb {

static int moments_groups(int dim) {
    int64_t g = (4ll << 20) / (dim > 0 ? dim : 1);
    if (g < 16) g = 16;
    if (g > 1024) g = 1024;
    return (int)g;
}

// partial[g][c] = (sum, sumsq) of (x - shift_c) over the rows owned by row-group g; shift_c = x[0][c].
__global__ void __launch_bounds__(256) moments_partial_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows,
                                                              int dim, int cpb, int rpi, double2* __restrict__ partial) {
    extern __shared__ double2 sm[];   // [rpi][cpb]
    const int tid = threadIdx.x;
    const int cl = tid % cpb, rl = tid / cpb;
    const int col = blockIdx.y * cpb + cl;
    double s = 0.0, ss = 0.0;
    const bool active = (rl < rpi) && (col < dim);
    if (active) {
        const double shift = (double)x[col];
        for (int64_t r = (int64_t)blockIdx.x * rpi + rl; r < rows; r += (int64_t)gridDim.x * rpi) {
            const double d = (double)x[r * ldx + col] - shift;
            s += d;
            ss += d * d;
        }
    }
    if (rl < rpi) sm[rl * cpb + cl] = make_double2(s, ss);
    __syncthreads();
    if (rl == 0 && col < dim) {
        for (int k = 1; k < rpi; ++k) {
            const double2 o = sm[k * cpb + cl];
            s += o.x;
            ss += o.y;
        }
        partial[(int64_t)blockIdx.x * dim + col] = make_double2(s, ss);
    }
}

// one warp per column: lanes stride over the row-group partials (fixed lane->group mapping + butterfly -> deterministic)
__global__ void moments_finalize_kernel(const float* __restrict__ x, const double2* __restrict__ partial, int groups,
                                        int64_t rows, int dim, float* __restrict__ mean, float* __restrict__ var) {
    const int lane = threadIdx.x & 31;
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (col >= dim) return;
    double s = 0.0, ss = 0.0;
    for (int g = lane; g < groups; g += 32) {
        const double2 p = partial[(int64_t)g * dim + col];
        s += p.x;
        ss += p.y;
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) {
        const double n = (double)rows;
        const double m = s / n;
        mean[col] = (float)((double)x[col] + m);
        // unbiased (torch.var default correction=1); rows == 1 gives NaN in torch as well
        var[col] = (float)((ss - s * m) / (n - 1.0));
    }
}

// running_mean_std.py:49-62, all in float64 like the TorchScript function.
__global__ void rms_merge_kernel(double* __restrict__ mean, double* __restrict__ var, double* __restrict__ count,
                                 const float* __restrict__ bmean, const float* __restrict__ bvar, double bcount, int dim) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const double cnt = count[0];
    if (c < dim) {
        const double delta = (double)bmean[c] - mean[c];
        const double tot = cnt + bcount;
        const double new_mean = mean[c] + delta * bcount / tot;
        const double m_a = var[c] * cnt;
        const double m_b = (double)bvar[c] * bcount;
        const double m2 = m_a + m_b + (delta * delta) * cnt * bcount / tot;
        mean[c] = new_mean;
        var[c] = m2 / tot;
    }
}
__global__ void rms_count_kernel(double* count, double bcount) { count[0] += bcount; }

}  // namespace sfb

using namespace sfb;

extern "C" {

int64_t sfb200_moments_workspace_bytes(int dim) { return (int64_t)moments_groups(dim) * dim * (int64_t)sizeof(double2); }

int sfb200_batch_moments(const float* x, int64_t ldx, int64_t rows, int dim, float* batch_mean, float* batch_var,
                         void* workspace, void* stream) {
    SFB_CHECK_ARG(x && batch_mean && batch_var && workspace && rows > 0 && dim > 0, "batch_moments: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int cpb = dim >= 256 ? 256 : dim;
    const int rpi = 256 / cpb;
    int groups = moments_groups(dim);
    const int64_t max_useful = ceil_div(rows, (int64_t)rpi * 4);
    if (groups > max_useful) groups = (int)(max_useful > 0 ? max_useful : 1);
    dim3 grid((unsigned)groups, (unsigned)ceil_div(dim, cpb));
    moments_partial_kernel<<<grid, 256, (size_t)rpi * cpb * sizeof(double2), st>>>(x, ldx, rows, dim, cpb, rpi,
                                                                                   (double2*)workspace);
    SFB_LAUNCH_OK();
    moments_finalize_kernel<<<(unsigned)ceil_div((int64_t)dim * 32, 128), 128, 0, st>>>(x, (const double2*)workspace, groups, rows, dim,
                                                                          batch_mean, batch_var);
    SFB_LAUNCH_OK();
    return 0;
}

int sfb200_rms_merge(double* mean, double* var, double* count, const float* batch_mean, const float* batch_var,
                     double batch_count, int dim, void* stream) {
    SFB_CHECK_ARG(mean && var && count && batch_mean && batch_var && dim > 0, "rms_merge: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    rms_merge_kernel<<<(unsigned)ceil_div(dim, 128), 128, 0, st>>>(mean, var, count, batch_mean, batch_var, batch_count,
                                                                   dim);
    SFB_LAUNCH_OK();
    rms_count_kernel<<<1, 1, 0, st>>>(count, batch_count);
    SFB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
