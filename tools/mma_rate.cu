// tcgen05.mma issue-rate probe (B200): cycles per instruction for kind::tf32 (K=8) and kind::f16 (K=16), A from shared
// memory (SS) or tensor memory (TS), N = 256 / 128 and the [wide | cross] pair the 3-pass split GEMMs issue.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I sample_factory_b200/csrc -I include tools/mma_rate.cu -o tools/bin/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace sfb;

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
                 "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {      // a/b format F16 = 0, c format F32 = 1
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) mma_rate(int mode, int iters, unsigned long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1024(smem_raw);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0h pairs / small fp32
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 32) {
        const uint64_t db = make_smem_desc(smem_u32(smem), false);            // up to 256 rows x 128 B
        const uint64_t da = make_smem_desc(smem_u32(smem + 32768), false);    // 128 rows x 128 B
        const uint32_t a_t = tm + 256;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint64_t bo = (uint64_t)((i & 3) * 2);
            const uint32_t ao = (uint32_t)((i & 3) * 8);
            switch (mode) {
                case 0: umma_tf32(tm, da + bo, db + bo, make_idesc(false, false, 128, 256), 1); break;
                case 1: umma_tf32_ts(tm, a_t + ao, db + bo, make_idesc(false, false, 128, 256), 1); break;
                case 2: umma_f16_ss(tm, da + bo, db + bo, idesc_f16(128, 256), 1); break;
                case 3: umma_f16_ts(tm, a_t + ao, db + bo, idesc_f16(128, 256), 1); break;
                case 4: umma_tf32_ts(tm, a_t + ao, db + bo, make_idesc(false, false, 128, 128), 1); break;
                case 5: umma_f16_ts(tm, a_t + ao, db + bo, idesc_f16(128, 128), 1); break;
                case 6:
                    umma_tf32_ts(tm, a_t + ao, db + bo, make_idesc(false, false, 128, 256), 1);
                    umma_tf32_ts(tm + 128, a_t + 32 + ao, db + bo, make_idesc(false, false, 128, 128), 1);
                    break;
                case 7:
                    umma_f16_ts(tm, a_t + ao, db + bo, idesc_f16(128, 256), 1);
                    umma_f16_ts(tm + 128, a_t + 32 + ao, db + bo, idesc_f16(128, 128), 1);
                    break;
                case 8: umma_f16_ss(tm, da + bo, db + bo, idesc_f16(128, 128), 1); break;
                case 9: umma_tf32(tm, da + bo, db + bo, make_idesc(false, false, 128, 128), 1); break;
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[mode] = (unsigned long long)(t1 - t0);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    unsigned long long* out;
    cudaMallocManaged(&out, 16 * sizeof(unsigned long long));
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200);
    const char* names[10] = {"tf32 SS N=256 K=8", "tf32 TS N=256 K=8", "f16  SS N=256 K=16", "f16  TS N=256 K=16", "tf32 TS N=128 K=8",
                             "f16  TS N=128 K=16", "tf32 TS pair 256+128 K=8", "f16  TS pair 256+128 K=16", "f16  SS N=128 K=16",
                             "tf32 SS N=128 K=8"};
    const int iters = 4096;
    for (int grid : {1, 148}) {
        printf("grid %d CTAs (one per SM)\n", grid);
        for (int mode = 0; mode < 10; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                mma_rate<<<grid, 128, 51200>>>(mode, iters, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            printf("  %-28s %8.1f cycles / iteration\n", names[mode], (double)out[mode] / iters);
        }
    }
    return 0;
}
