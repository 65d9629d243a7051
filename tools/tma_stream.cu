// Streaming-floor microbenchmark: how fast can 148 persistent CTAs pull the trellis through the TMA ring used by the
// decode-GEMM kernels when nothing else happens?  Not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I exllamav3_b200/csrc -o tools/tma_stream tools/tma_stream.cu
//   tools/tma_stream            (prints a table: K, stages, consumer warps, read-back, GB/s)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

using namespace exl3b::ptx;

struct P { int KB, strips, S, NC, read, wbytes, boxcols; uint32_t* sink; };

__global__ void __launch_bounds__(1024, 1) stream_kernel(const P p, const __grid_constant__ CUtensorMap tm)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * p.wbytes);
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (S + s); };
    if (threadIdx.x == 0)
    {
        for (int s = 0; s < S; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), p.NC); }
        fence_barrier_init();
    }
    __syncthreads();
    const long long U = (long long) p.KB * p.strips;
    const int G = gridDim.x;
    const long long ubeg = U * blockIdx.x / G, uend = U * (blockIdx.x + 1) / G;
    const int n_units = (int) (uend - ubeg);
    if (warp == 0)
    {
        const uint64_t pol = policy_evict_first();
        int strip = (int) (ubeg / p.KB), kb = (int) (ubeg % p.KB);
        int s = 0, ph = 0;
        for (int u = 0; u < n_units; ++u)
        {
            if (u >= S) mbar_wait<64>(EMPTY(s), ph ^ 1);
            if (elect_one())
            {
                mbar_arrive_expect_tx(FULL(s), (uint32_t) p.wbytes);
                tma_load_2d(smem_u32(smem) + s * p.wbytes, &tm, strip * p.boxcols, kb * 8, FULL(s), pol);
            }
            if (++kb == p.KB) { kb = 0; ++strip; }
            if (++s == S) { s = 0; ph ^= 1; }
        }
    }
    else if (warp <= p.NC)
    {
        int s = 0, ph = 0;
        uint32_t acc = 0;
        const int cw = warp - 1;
        for (int u = 0; u < n_units; ++u)
        {
            mbar_wait<32>(FULL(s), ph);
            if (p.read)
            {
                const uint4* src = reinterpret_cast<const uint4*>(smem + s * p.wbytes);
                for (int i = cw * 32 + lane; i < p.wbytes / 16; i += p.NC * 32) { uint4 v = src[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(EMPTY(s));
            if (++s == S) { s = 0; ph ^= 1; }
        }
        if (acc == 0x12345u) p.sink[0] = acc;
    }
}

__global__ void ldg_kernel(const uint4* __restrict__ src, size_t n16, uint32_t* sink)
{
    uint32_t acc = 0;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t) gridDim.x * blockDim.x)
    { uint4 v = __ldcs(src + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345u) sink[0] = acc;
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int k = 4096;
    PFN_encode encode; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**) &encode, cudaEnableDefault, &q);
    uint32_t* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int n, K; };
    const Cfg shapes[] = { {14336, 4}, {128256 / 128 * 128, 6}, {4096, 4} };
    for (const Cfg& sh : shapes)
    {
        const int n = sh.n, K = sh.K;
        const size_t bytes = (size_t) k * n * K / 8;
        const int reps = bytes < (64u << 20) ? 16 : 1;       // several distinct copies so that every launch reads cold data
        uint8_t* buf; cudaMalloc(&buf, bytes * reps); cudaMemset(buf, 1, bytes * reps);
        {
            float best = 1e9;
            for (int it = 0; it < 5; ++it)
            {
                cudaEventRecord(e0);
                ldg_kernel<<<148 * 8, 512>>>((const uint4*) buf, bytes * reps / 16, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("shape k=%d n=%d K=%d  (%.1f MB)  plain LDG read: %.0f GB/s\n", k, n, K, bytes / 1e6, bytes * reps / best / 1e6);
        }
        for (int boxrows_mult = 1; boxrows_mult <= 1; ++boxrows_mult)
        for (int S : {4, 8, 16, 24})
        for (int NC : {1, 8, 17})
        for (int rd : {0, 1})
        {
            P p; p.KB = k / 128; p.strips = n / 128; p.S = S; p.NC = NC; p.read = rd; p.wbytes = 2048 * K; p.boxcols = 32 * K; p.sink = sink;
            const int smem = S * p.wbytes + 16 * S + 64;
            if (smem > 220 * 1024) continue;
            std::vector<CUtensorMap> maps(reps);
            for (int r = 0; r < reps; ++r)
            {
                cuuint64_t row_bytes = (cuuint64_t) (n / 16) * 32 * K;
                cuuint64_t gdim[2] = { row_bytes / 8, (cuuint64_t) (k / 16) };
                cuuint64_t gstride[1] = { row_bytes };
                cuuint32_t box[2] = { (cuuint32_t) (32 * K), 8 };
                cuuint32_t estr[2] = { 1, 1 };
                CUresult cr = encode(&maps[r], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, buf + bytes * r, gdim, gstride, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int) cr); return 1; }
            }
            float best = 1e9;
            for (int it = 0; it < 4; ++it)
            {
                cudaEventRecord(e0);
                for (int r = 0; r < reps; ++r) stream_kernel<<<148, 32 * (1 + NC), smem>>>(p, maps[r]);
                cudaEventRecord(e1);
                if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("  S=%2d consumers=%2d read=%d : %7.1f us/launch  %6.0f GB/s\n", S, NC, rd, best * 1e3 / reps, bytes * reps / best / 1e6);
        }
        cudaFree(buf);
    }
    return 0;
}
