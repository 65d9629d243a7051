// tcgen05.mma issue/execute rate for the small-N shapes of the swap-AB decode-GEMM (A operand in TMEM).  Not product code.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I exllamav3_b200/csrc -o tools/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
using namespace exl3b::ptx;

// mode 0: kind::i8 A=tmem, 1: kind::f16 A=tmem, 2: kind::f16 A=smem
// 16 i8 MMAs followed by n_commits commits (to distinct barriers); the issuer waits only every `wait_every` groups on the
// last barrier: how much serial time does a commit add to the issuing warp / the tensor pipe?
__global__ void __launch_bounds__(128, 1) commit_kernel(int n_commits, int wait_every, int iters, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    if (warp == 0)
    {
        const uint32_t idesc = idesc_u8s8_s32(128, 16);
        const uint64_t bdesc = smem_desc(smem_u32(smem), 128, 4096, 0);
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
        {
            if (elect_one())
            {
                #pragma unroll
                for (int j = 0; j < 16; ++j) mma_i8_ts(tb + 384, tb + 8 * j, bdesc + 16 * j, idesc, 1);
                for (int c = 0; c < n_commits; ++c) tc_commit(smem_u32(&bars[c]));
            }
            __syncwarp();
            if ((it + 1) % wait_every == 0)
            {
                // every barrier got wait_every arrivals of count 1: phases advance wait_every times; parity tracking only
                // works for wait_every == 1, so for the pipelined variant the barriers are simply not waited on except the last
                mbar_wait(smem_u32(&bars[n_commits - 1]), ph); ph ^= 1;
            }
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int per_commit, int iters, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    if (warp == 0)
    {
        const uint32_t idesc = MODE == 0 ? idesc_u8s8_s32(128, N) : idesc_f16_f32(128, N);
        const uint64_t bdesc = smem_desc(smem_u32(smem), 128, MODE == 0 ? 4096 : 2048, 0);
        const uint64_t adesc = smem_desc(smem_u32(smem) + 65536, 128, 2048, 0);
        const uint32_t d_addr = tb + 384;
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
        {
            if (elect_one())
            {
                for (int j = 0; j < per_commit; ++j)
                {
                    if (MODE == 0) mma_i8_ts(d_addr, tb + 8 * (j & 15), bdesc + 16 * (j & 15), idesc, 1);
                    if (MODE == 1) mma_f16_ts(d_addr, tb + 8 * (j & 15), bdesc + 16 * (j & 15), idesc, 1);
                    if (MODE == 2) mma_f16_ss(d_addr, adesc + 16 * (j & 7), bdesc + 16 * (j & 7), idesc, 1);
                }
                tc_commit(smem_u32(&bar));
            }
            __syncwarp();
            mbar_wait(smem_u32(&bar), ph); ph ^= 1;
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main()
{
    long long* out; cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int nc : {1, 2, 3, 4})
    {
        for (int rep = 0; rep < 2; ++rep) commit_kernel<<<148, 128, 160 * 1024>>>(nc, 1, 512, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("commit_kernel: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("16 i8 MMAs + %d commits, wait each group: %7.1f cycles per group\n", nc, (double) h[0] / 512);
    }
    const char* names[] = { "kind::i8  A=tmem (K=32 B)", "kind::f16 A=tmem (K=16)", "kind::f16 A=smem (K=16)" };
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {16, 128})
            for (int per : {16, 128})
            {
                if (mode != 0 && N == 8) continue;
                const int iters = per == 16 ? 512 : 64;
                for (int rep = 0; rep < 2; ++rep)
                {
                    if (mode == 0) rate_kernel<0><<<148, 128, 160 * 1024>>>(N, per, iters, out);
                    if (mode == 1) rate_kernel<1><<<148, 128, 160 * 1024>>>(N, per, iters, out);
                    if (mode == 2) rate_kernel<2><<<148, 128, 160 * 1024>>>(N, per, iters, out);
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
                long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                printf("%-28s N=%3d  %3d MMAs/commit: %7.1f cycles per MMA  (%.0f cycles per commit group)\n", names[mode], N, per,
                       (double) h[0] / iters / per, (double) h[0] / iters);
            }
    return 0;
}
