// Data-parallel exchange over NVLink peer memory (the reference has no collective at all).
//
// One process per GPU.  Every rank owns ONE "comm buffer" in its HBM that all peers map through CUDA IPC:
//
//     [0, 8 KiB)          flags[256 blocks][8 source ranks]   uint32, written by PEERS (st.release.sys), spun on locally
//     [8 KiB, 16 KiB)     local control words (sequence number, block ticket, grid barrier) -- never touched by peers
//     [16 KiB, +scratch)  float64 scratch: the rank's contribution to a small all-reduce (moments, loss statistics)
//     [.., +4*n)          the rank's flat fp32 gradient -- the backward kernels write it HERE, so there is no staging copy
//
// All exchanges are ONE-SHOT pulls: barrier -> every rank reads every peer's contribution over NVLink and sums it in
// rank order 0..G-1 (fixed order => bit-identical results on all ranks, which keeps the replicas identical) -> barrier
// (nobody overwrites a buffer a peer may still be reading).  At 1.2 MB per gradient this is the latency regime: the
// kernel costs two NVLink flag round trips plus (G-1) x 1.2 MB of peer reads (~2 us at the measured 770 GB/s).
//
// sfb200_dp_grad_allreduce_clip_adam: the gradient all-reduce, the global grad-norm, the clip and the Adam update are ONE
// kernel: phase 1 pulls + sums + writes the reduced gradient to a local buffer and leaves per-block sums of squares, a
// device-wide barrier (all blocks co-resident: grid <= #SMs) makes the norm available, phase 2 is clip_adam_body.
// Nothing here allocates or synchronises the host; the sequence number lives in device memory, so the launches can be
// captured in a CUDA graph and replayed (this is what lets the data-parallel learner run as one graph).
#include <cuda.h>
#include <string.h>

#include "adam_core.cuh"
#include "common.cuh"

namespace sfb {

constexpr int kDpMaxWorld = 8;
constexpr int kDpMaxBlocks = 256;
constexpr int64_t kDpFlagBytes = 8192;
constexpr int64_t kDpHeaderBytes = 16384;

struct DpComm {
    int rank, world;
    uint8_t* base[kDpMaxWorld];   // base[rank] is the local buffer
    int64_t scratch_bytes;
};

struct DpCtl {   // local control words at base[rank] + kDpFlagBytes
    unsigned long long seq;      // barriers completed so far (every launch adds 2)
    unsigned int ticket;         // blocks that have finished the current launch
    unsigned int grid_bar;       // blocks that have reached the mid-kernel device barrier
};

static DpComm g_comms[16];
static int g_comm_used[16] = {0};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ DpCtl* dp_ctl(const DpComm& c) { return reinterpret_cast<DpCtl*>(c.base[c.rank] + kDpFlagBytes); }

// Block-level barrier across ranks: block b of every rank meets block b of every other rank.  Whatever the threads of
// this block wrote (or finished reading) before the call is ordered before the peers' accesses after it.
__device__ __forceinline__ void dp_block_barrier(const DpComm& c, uint32_t seq) {
    __syncthreads();
    if ((int)threadIdx.x < c.world) {
        const int peer = threadIdx.x;
        uint32_t* dst = reinterpret_cast<uint32_t*>(c.base[peer]) + blockIdx.x * kDpMaxWorld + c.rank;
        st_release_sys(dst, seq);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(c.base[c.rank]) + blockIdx.x * kDpMaxWorld + peer;
        while ((int32_t)(ld_acquire_sys(src) - seq) < 0) {
        }
    }
    __syncthreads();
}

// first thing in every comm kernel: the sequence number of this launch (identical on all ranks: same launch sequence)
__device__ __forceinline__ uint32_t dp_begin(const DpComm& c) {
    __shared__ uint32_t s_seq;
    if (threadIdx.x == 0) s_seq = (uint32_t)*reinterpret_cast<volatile unsigned long long*>(&dp_ctl(c)->seq);
    __syncthreads();
    return s_seq;
}
// last thing: the block that finishes last advances the sequence number (every block has read it by then)
__device__ __forceinline__ void dp_end(const DpComm& c, uint32_t seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
        DpCtl* ctl = dp_ctl(c);
        __threadfence();
        if (atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1u) {
            ctl->ticket = 0u;
            ctl->grid_bar = 0u;
            ctl->seq = (unsigned long long)seq + 2ull;
            __threadfence();
        }
    }
}

__device__ __forceinline__ double block_sum_256(double v) {
    __shared__ double sm[8];
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w];
    return t;
}

// ---------------------------------------------------------------------------------------------- gradient all-reduce
template <bool FUSE_ADAM>
__global__ void __launch_bounds__(256) dp_grad_allreduce_kernel(const DpComm c, int64_t grad_off, float* __restrict__ g_out,
                                                                int64_t n, AdamArgs adam, double* __restrict__ part) {
    const uint32_t seq = dp_begin(c);
    dp_block_barrier(c, seq + 1u);   // every rank's backward has finished writing its gradient
    const float* src[kDpMaxWorld];
#pragma unroll
    for (int r = 0; r < kDpMaxWorld; ++r) src[r] = r < c.world ? reinterpret_cast<const float*>(c.base[r] + grad_off) : nullptr;
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double ss = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 acc = reinterpret_cast<const float4*>(src[0])[i];
#pragma unroll
        for (int r = 1; r < kDpMaxWorld; ++r) {
            if (r < c.world) {
                const float4 v = reinterpret_cast<const float4*>(src[r])[i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(g_out)[i] = acc;
        ss += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {   // tail (< 4 elements)
        float acc = src[0][i];
        for (int r = 1; r < c.world; ++r) acc += src[r][i];
        g_out[i] = acc;
        ss += (double)acc * acc;
    }
    ss = block_sum_256(ss);
    if (threadIdx.x == 0) part[blockIdx.x] = ss;
    __threadfence();   // phase 2 reads the reduced gradient / the partials written by OTHER blocks
    dp_block_barrier(c, seq + 2u);   // every rank has finished reading this block's share of every gradient
    if (FUSE_ADAM) {
        // device-wide barrier (grid <= #SMs, all blocks resident): the partial sums of squares of ALL blocks are in `part`
        DpCtl* ctl = dp_ctl(c);
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&ctl->grid_bar, 1u);
            while (ld_acquire_gpu(&ctl->grid_bar) < gridDim.x) {
            }
        }
        __syncthreads();
        clip_adam_body(adam, part, (int)gridDim.x);
    }
    dp_end(c, seq);
}

// ---------------------------------------------------------------------------------------------- small fp64 all-reduce
// buf[0..n): element i belongs to column i % row_len; per column the combine is SUM, MAX (max_mask), MIN (min_mask), MEAN
// over the ranks (avg_mask) or KEEP (keep_mask: already identical on all ranks, left untouched).  row_len = 0: all summed.
// MOMENTS mode (running_mean_std.py:72-77 made global): the contribution is built from per-rank (mean, unbiased var) over
// `rows` rows as [sum x | sum x^2] and the pooled (mean, unbiased var) of all G*rows rows is written back as fp32.
template <bool MOMENTS>
__global__ void __launch_bounds__(256) dp_allreduce_f64_kernel(const DpComm c, double* __restrict__ buf, int n, int row_len,
                                                               unsigned long long max_mask, unsigned long long min_mask,
                                                               unsigned long long keep_mask, unsigned long long avg_mask,
                                                               float* __restrict__ bmean,
                                                               float* __restrict__ bvar, double rows) {
    const uint32_t seq = dp_begin(c);
    double* mine = reinterpret_cast<double*>(c.base[c.rank] + kDpHeaderBytes);
    const int total = MOMENTS ? 2 * n : n;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (MOMENTS) {
            const int j = i < n ? i : i - n;
            const double m = (double)bmean[j];
            mine[i] = i < n ? m * rows : (double)bvar[j] * (rows - 1.0) + m * m * rows;
        } else {
            mine[i] = buf[i];
        }
    }
    dp_block_barrier(c, seq + 1u);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (MOMENTS ? n : total); i += stride) {
        if (MOMENTS) {
            double s0 = 0.0, s1 = 0.0;
            for (int r = 0; r < c.world; ++r) {
                const double* p = reinterpret_cast<const double*>(c.base[r] + kDpHeaderBytes);
                s0 += p[i];
                s1 += p[n + i];
            }
            const double tot = rows * (double)c.world;
            const double gmean = s0 / tot;
            const double gm2 = s1 - gmean * gmean * tot;
            bmean[i] = (float)gmean;
            bvar[i] = (float)(gm2 / (tot - 1.0));
        } else {
            const int col = row_len > 0 ? i % row_len : 0;
            const unsigned long long bit = row_len > 0 ? (1ull << col) : 0ull;
            if (bit & keep_mask) continue;
            double acc = reinterpret_cast<const double*>(c.base[0] + kDpHeaderBytes)[i];
            for (int r = 1; r < c.world; ++r) {
                const double v = reinterpret_cast<const double*>(c.base[r] + kDpHeaderBytes)[i];
                acc = (bit & max_mask) ? fmax(acc, v) : ((bit & min_mask) ? fmin(acc, v) : acc + v);
            }
            buf[i] = (bit & avg_mask) ? acc / (double)c.world : acc;
        }
    }
    dp_block_barrier(c, seq + 2u);
    dp_end(c, seq);
}

// out[0] = sum over rows of src[row * stride + col]  (global valid count from the all-reduced minibatch partials)
__global__ void colsum_f64_kernel(const double* __restrict__ src, int rows, int stride, int col, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int r = 0; r < rows; ++r) s += src[(int64_t)r * stride + col];
        out[0] = s;
    }
}

static const DpComm* get_comm(int comm) {
    if (comm < 0 || comm >= 16 || !g_comm_used[comm]) return nullptr;
    return &g_comms[comm];
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int sfb200_ipc_export(const void* ptr, void* handle_out_host, int64_t* offset_out_host) {
    SFB_CHECK_ARG(ptr && handle_out_host && offset_out_host, "ipc_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    // (the driver API is resolved at run time: the library must load on machines without libcuda, e.g. the CPU test box)
    typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        set_error("ipc_export: cuMemGetAddressRange is not available");
        return 2;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (reinterpret_cast<GetRangeFn>(fn)(&base, &size, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) {
        set_error("ipc_export: cuMemGetAddressRange failed (not a cudaMalloc allocation?)");
        return 2;
    }
    cudaIpcMemHandle_t h;
    SFB_CUDA_OK(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
    memcpy(handle_out_host, &h, sizeof(h));
    *offset_out_host = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
    return 0;
}

int sfb200_ipc_import(const void* handle_host, int64_t offset, void** ptr_out_host) {
    SFB_CHECK_ARG(handle_host && ptr_out_host && offset >= 0, "ipc_import: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void* base = nullptr;
    SFB_CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out_host = (void*)((uint8_t*)base + offset);
    return 0;
}

int sfb200_ipc_close(void* ptr, int64_t offset) {
    if (!ptr) return 0;
    SFB_CUDA_OK(cudaIpcCloseMemHandle((void*)((uint8_t*)ptr - offset)));
    return 0;
}

int64_t sfb200_dp_header_bytes(void) { return kDpHeaderBytes; }

int sfb200_dp_create(int rank, int world, const uint64_t* peer_ptrs_host, int64_t scratch_bytes) {
    if (!(rank >= 0 && world >= 2 && world <= kDpMaxWorld && rank < world && peer_ptrs_host && scratch_bytes >= 4096)) {
        set_error("dp_create: bad arguments (2 <= world <= 8)");
        return -1;
    }
    for (int i = 0; i < 16; ++i) {
        if (!g_comm_used[i]) {
            DpComm& c = g_comms[i];
            c.rank = rank;
            c.world = world;
            c.scratch_bytes = scratch_bytes;
            for (int r = 0; r < kDpMaxWorld; ++r) c.base[r] = r < world ? (uint8_t*)(uintptr_t)peer_ptrs_host[r] : nullptr;
            g_comm_used[i] = 1;
            return i;
        }
    }
    set_error("dp_create: communicator table full");
    return -1;
}

int sfb200_dp_destroy(int comm) {
    if (comm >= 0 && comm < 16) g_comm_used[comm] = 0;
    return 0;
}

static int dp_grad_impl(int comm, float* g_out, int64_t n, bool fuse, const AdamArgs& aa, void* workspace, void* stream) {
    const DpComm* c = get_comm(comm);
    SFB_CHECK_ARG(c && g_out && n > 0 && workspace, "dp_grad_allreduce: bad arguments");
    SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(g_out) & 15u) == 0, "dp_grad_allreduce: g_out must be 16-byte aligned");
    const int64_t grad_off = kDpHeaderBytes + c->scratch_bytes;
    int64_t blocks = ceil_div(ceil_div(n, 4), 256);
    int64_t cap = sm_count();
    if (cap > kDpMaxBlocks) cap = kDpMaxBlocks;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (fuse) dp_grad_allreduce_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(*c, grad_off, g_out, n, aa, (double*)workspace);
    else dp_grad_allreduce_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(*c, grad_off, g_out, n, aa, (double*)workspace);
    SFB_LAUNCH_OK();
    return 0;
}

int sfb200_dp_grad_allreduce(int comm, float* g_out, int64_t n, void* workspace, void* stream) {
    AdamArgs none = {};
    return dp_grad_impl(comm, g_out, n, false, none, workspace, stream);
}

int sfb200_dp_grad_allreduce_clip_adam(int comm, float* g_out, float* p, float* m, float* v, int64_t n, int64_t step,
                                       const int64_t* steps_done_dev, double lr, const double* lr_dev, double beta1,
                                       double beta2, double eps, double max_grad_norm, const double* lr_scale_num,
                                       const double* lr_scale_den, float* grad_norm_out, void* workspace, void* stream) {
    SFB_CHECK_ARG(p && m && v && (step >= 1 || steps_done_dev), "dp_grad_allreduce_clip_adam: bad arguments");
    SFB_CHECK_ARG((lr_scale_num == nullptr) == (lr_scale_den == nullptr), "dp_grad_allreduce_clip_adam: lr_scale num/den mismatch");
    const AdamArgs aa = make_adam_args(p, g_out, m, v, n, lr, lr_dev, beta1, beta2, step, steps_done_dev, eps, max_grad_norm,
                                       lr_scale_num, lr_scale_den, grad_norm_out);
    return dp_grad_impl(comm, g_out, n, true, aa, workspace, stream);
}

int sfb200_dp_allreduce_f64(int comm, double* buf, int n, int row_len, uint64_t max_mask, uint64_t min_mask,
                            uint64_t keep_mask, uint64_t avg_mask, void* stream) {
    const DpComm* c = get_comm(comm);
    SFB_CHECK_ARG(c && buf && n > 0 && row_len >= 0 && row_len <= 64, "dp_allreduce_f64: bad arguments");
    SFB_CHECK_ARG((int64_t)n * 8 <= c->scratch_bytes, "dp_allreduce_f64: %d doubles exceed the communicator's scratch", n);
    int blocks = (int)ceil_div(n, 256);
    if (blocks > 64) blocks = 64;
    dp_allreduce_f64_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(*c, buf, n, row_len, max_mask, min_mask, keep_mask,
                                                                          avg_mask, nullptr, nullptr, 0.0);
    SFB_LAUNCH_OK();
    return 0;
}

int sfb200_dp_pooled_moments(int comm, float* batch_mean, float* batch_var, int dim, double rows_per_rank, void* stream) {
    const DpComm* c = get_comm(comm);
    SFB_CHECK_ARG(c && batch_mean && batch_var && dim > 0 && rows_per_rank >= 1.0, "dp_pooled_moments: bad arguments");
    SFB_CHECK_ARG((int64_t)dim * 16 <= c->scratch_bytes, "dp_pooled_moments: dim %d exceeds the communicator's scratch", dim);
    int blocks = (int)ceil_div(dim, 256);
    if (blocks > 64) blocks = 64;
    dp_allreduce_f64_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(*c, nullptr, dim, 0, 0ull, 0ull, 0ull, 0ull, batch_mean,
                                                                         batch_var, rows_per_rank);
    SFB_LAUNCH_OK();
    return 0;
}

int sfb200_colsum_f64(const double* src, int rows, int stride, int col, double* out, void* stream) {
    SFB_CHECK_ARG(src && out && rows > 0 && stride > 0 && col >= 0 && col < stride, "colsum_f64: bad arguments");
    colsum_f64_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(src, rows, stride, col, out);
    SFB_LAUNCH_OK();
    return 0;
}

}  // extern "C"
