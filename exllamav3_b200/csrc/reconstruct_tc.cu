// reconstruct_had on the tensor cores:  W = diag(suh) . H128 . W_hat . H128 . diag(svh)  per 128 x 128 block, both 128-point
// Hadamards as tcgen05 GEMMs against the +-1 matrix H128 (exact in fp16, fp32 accumulation) instead of 2 x 7 butterfly
// stages on the CUDA cores (kernels_basic.cu: 128 fp32 registers per thread, two CTAs per SM, 45 us for 4096 x 4096
// where the bytes need ~7; this kernel: 25 us, profiles/r02_reconstruct_had_tc_splits.jsonl).  Replaces the reference's reconstruct_had_kernel (exllamav3_ext/quant/reconstruct.cu:159-306).
//
// Per block (one CTA of 128 x SPLIT threads -- SPLIT threads share a column / row, default 2 --, two CTAs per SM, persistent over
// the blocks; the next block's packed words are fetched one block ahead):
//   1. thread = one column n: decode its 128 k-values (8 tiles, decode16) and store them as the B operand W_hat^T [n][k]
//      (K-major core-matrix layout) in shared memory
//   2. MMA 1 (8 x UTCHMMA, M = N = 128):  D[k'][n] = sum_k H[k'][k] W_hat[k][n]         (A = H from shared memory)
//   3. thread = one row k': tcgen05.ld its row, * 1/sqrt(128), round to fp16 (the reference keeps an fp16 tile between
//      the passes), tcgen05.st as the A operand of the second GEMM (A from TMEM)
//   4. MMA 2:  D[k'][n'] = sum_n T[k'][n] H[n][n']                                      (B = the same H)
//   5. thread = one row: tcgen05.ld, * 1/sqrt(128) -> fp16, * suh[row] * svh[col] in fp16 (the rounding order of
//      reconstruct.cu:300-304), 256 contiguous bytes per row
// H128 is written to shared memory once per CTA; TMEM: 128 accumulator columns (reused by both GEMMs) + 64 operand columns.
#include "tc_common.cuh"
#include <cstdlib>

namespace exl3b {

using namespace ptx;

constexpr int RT_THREADS = 128;
constexpr int RT_H_BYTES = 128 * 256;                  // H128 fp16, K-major core-matrix layout
constexpr int RT_B_BYTES = 128 * 256;                  // W_hat^T block
constexpr int RT_SMEM = RT_H_BYTES + RT_B_BYTES + 256 + 64;
constexpr int RT_DEFAULT_SPLIT = 2;       // measured (profiles/r02_reconstruct_had_tc_splits.jsonl): 2 threads per row is the fastest

// (row, k) of a 128 x 128 fp16 operand in the K-major no-swizzle core-matrix layout: 8 rows x 16 B core matrices,
// K-adjacent ones 128 B apart (LBO), 8-row groups 2048 B apart (SBO)
__device__ __forceinline__ uint32_t rt_off(int row, int k) { return (row >> 3) * 2048 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2; }

__device__ __forceinline__ void tmem_ld_32x32b_x32_rt(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// SPLIT = 1, 2 or 4 warps per TMEM lane quarter: the CTA has 128 * SPLIT threads; the SPLIT threads of a column / row share its
// decode (8 / SPLIT tiles each) and its read-outs (128 / SPLIT accumulator columns each).  More threads per block = shorter
// serial chain per block and more warps per SM to hide the latencies (the kernel is a chain of short dependent phases).
template <int K, int cb, int SPLIT>
__global__ void __launch_bounds__(RT_THREADS * SPLIT, 2)
reconstruct_had_tc_kernel(half* __restrict__ out, const uint32_t* __restrict__ packed, const half* __restrict__ suh,
                          const half* __restrict__ svh, int n_out, int packed_tiles_n, int tile_n_offset, int blocks_n, int n_blocks)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sH = smem;
    uint8_t* sB = smem + RT_H_BYTES;
    half* s_svh = reinterpret_cast<half*>(smem + RT_H_BYTES + RT_B_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RT_H_BYTES + RT_B_BYTES + 256);
    const uint32_t bar = smem_u32(bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, part = warp >> 2;                 // TMEM lane quarter, share of the column / row
    const int row = q * 32 + lane;                            // the thread's column in phase 1, its row in phases 3 and 5
    constexpr int COLS = 128 / SPLIT;                         // accumulator columns per thread and read-out

    // H128[row][c] = (-1)^popcount(row & c): 16 x 16 bytes per row, shared by the row's threads
    #pragma unroll
    for (int kc = part; kc < 16; kc += SPLIT)
    {
        uint32_t wv[4];
        #pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const int c0 = kc * 8 + 2 * j;
            const uint32_t lo = (__popc(row & c0) & 1) ? 0xbc00u : 0x3c00u;
            const uint32_t hi = (__popc(row & (c0 + 1)) & 1) ? 0xbc00u : 0x3c00u;
            wv[j] = lo | (hi << 16);
        }
        *reinterpret_cast<uint4*>(sH + rt_off(row, kc * 8)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<256>(smem_u32(tmem_slot));
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_base = (uint32_t) (q * 32) << 16;
    const uint32_t idesc = idesc_f16_f32(128, 128);
    const uint64_t desc_t = smem_desc(0, 128, 2048, 0);
    const uint32_t h_addr = smem_u32(sH), b_addr = smem_u32(sB);
    constexpr int D_COL = 0, A_COL = 128;

    const int tl = strip_tile(q, lane), col = strip_col(q, lane);
    int ph = 0;
    // the packed words, suh and svh of a block are fetched one block AHEAD (registers), right after the previous block's decode:
    // the DRAM latency then runs under the two GEMMs and the read-outs instead of opening the block's serial chain
    uint32_t w[8 / SPLIT][K + 1];
    half nx_suh, nx_svh = __float2half(0.f);
    auto fetch = [&](int blk)
    {
        const int kb = blk / blocks_n, nb = blk - kb * blocks_n;
        #pragma unroll
        for (int i = 0; i < 8 / SPLIT; ++i)
        {
            const int t = part + i * SPLIT;
            load_chunk<K>(packed + ((size_t) (kb * 8 + t) * packed_tiles_n + tile_n_offset + nb * 8 + tl) * (8 * K), lane & 7, w[i]);
        }
        nx_suh = suh[kb * 128 + row];
        if (threadIdx.x < 128) nx_svh = svh[nb * 128 + threadIdx.x];
    };
    if ((int) blockIdx.x < n_blocks) fetch(blockIdx.x);
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x)
    {
        const int kb = blk / blocks_n, nb = blk - kb * blocks_n;
        // ---- 1. decode this thread's share of its column into the B operand ----
        #pragma unroll
        for (int i = 0; i < 8 / SPLIT; ++i)
        {
            const int t = part + i * SPLIT;
            uint32_t o[8];
            if (q & 1) decode16<K, cb, 1>(w[i], o); else decode16<K, cb, 0>(w[i], o);
            *reinterpret_cast<uint4*>(sB + rt_off(col, 16 * t)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(sB + rt_off(col, 16 * t + 8)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        if (threadIdx.x < 128) s_svh[threadIdx.x] = nx_svh;
        const half my_suh = nx_suh;
        if (blk + (int) gridDim.x < n_blocks) fetch(blk + gridDim.x);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();                                       // B complete; every thread's tcgen05.ld of the previous block is done
        // ---- 2. left Hadamard: D = H x W_hat ----
        if (warp == 0)
        {
            tc_fence_after();
            if (elect_one())
            {
                #pragma unroll
                for (int j = 0; j < 8; ++j)
                    mma_f16_ss(tmem_base + D_COL, desc_t | (uint64_t) (((h_addr + j * 256) >> 4) & 0x3fff),
                               desc_t | (uint64_t) (((b_addr + j * 256) >> 4) & 0x3fff), idesc, j > 0);
                tc_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph); ph ^= 1;
        tc_fence_after();
        // ---- 3. this thread's share of its row of H.W_hat -> * 1/sqrt(128) -> fp16 -> A operand in TMEM ----
        #pragma unroll
        for (int c = 0; c < COLS / 32; ++c)
        {
            uint32_t r[32];
            tmem_ld_32x32b_x32_rt(tmem_base + lane_base + D_COL + part * COLS + c * 32, r);
            tc_wait_ld();
            uint32_t pk[16];
            #pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                const half2 h = __floats2half2_rn(__uint_as_float(r[2 * j]) * R_SCALE, __uint_as_float(r[2 * j + 1]) * R_SCALE);
                pk[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            tmem_st_32x32b_x16(tmem_base + lane_base + A_COL + part * (COLS / 2) + c * 16, pk);
        }
        tc_wait_st();
        tc_fence_before();
        __syncthreads();
        // ---- 4. right Hadamard: D = T x H (A from TMEM) ----
        if (warp == 0)
        {
            tc_fence_after();
            if (elect_one())
            {
                #pragma unroll
                for (int j = 0; j < 8; ++j)
                    mma_f16_ts(tmem_base + D_COL, tmem_base + A_COL + 8 * j, desc_t | (uint64_t) (((h_addr + j * 256) >> 4) & 0x3fff), idesc, j > 0);
                tc_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph); ph ^= 1;
        tc_fence_after();
        // ---- 5. scale and store this thread's share of its row: (fp16(h / sqrt(128)) * suh[row]) * svh[col], fp16 multiplies ----
        half* orow = out + ((size_t) kb * 128 + row) * n_out + nb * 128 + part * COLS;
        const half2 su2 = __half2half2(my_suh);
        #pragma unroll
        for (int c = 0; c < COLS / 32; ++c)
        {
            uint32_t r[32];
            tmem_ld_32x32b_x32_rt(tmem_base + lane_base + D_COL + part * COLS + c * 32, r);
            tc_wait_ld();
            #pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
            {
                uint32_t ov[4];
                #pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const int e = j4 * 8 + j * 2;
                    half2 h = __floats2half2_rn(__uint_as_float(r[e]) * R_SCALE, __uint_as_float(r[e + 1]) * R_SCALE);
                    h = __hmul2(__hmul2(h, su2), *reinterpret_cast<const half2*>(s_svh + part * COLS + c * 32 + e));
                    ov[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(orow + c * 32 + j4 * 8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
            }
        }
        // s_svh and the accumulator are rewritten by the next block: every thread must be through with them
        tc_fence_before();
        __syncthreads();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
    {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

static int g_rt_split = 0;          // 0 = default

template <int K, int cb, int SPLIT>
static void reconstruct_had_tc_launch_s(cudaStream_t stream, half* out, const uint16_t* packed, const half* suh, const half* svh,
                                        int k, int n_out, int packed_tiles_n, int tile_n_offset, int num_sms, cudaError_t* err)
{
    static bool attr_set[32] = {};
    int dev = 0; cudaGetDevice(&dev);
    if (!attr_set[dev & 31])
    {
        *err = cudaFuncSetAttribute(reconstruct_had_tc_kernel<K, cb, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM);
        if (*err != cudaSuccess) return;
        attr_set[dev & 31] = true;
    }
    const int blocks_n = n_out / 128, n_blocks = blocks_n * (k / 128);
    int grid = 2 * num_sms; if (grid > n_blocks) grid = n_blocks;
    reconstruct_had_tc_kernel<K, cb, SPLIT><<<grid, RT_THREADS * SPLIT, RT_SMEM, stream>>>(out, (const uint32_t*) packed, suh, svh, n_out,
                                                                                          packed_tiles_n, tile_n_offset, blocks_n, n_blocks);
    *err = cudaSuccess;
}

template <int K, int cb>
static void reconstruct_had_tc_launch(cudaStream_t stream, half* out, const uint16_t* packed, const half* suh, const half* svh,
                                      int k, int n_out, int packed_tiles_n, int tile_n_offset, int num_sms, cudaError_t* err)
{
    const int split = g_rt_split ? g_rt_split : RT_DEFAULT_SPLIT;
    if (split == 1) reconstruct_had_tc_launch_s<K, cb, 1>(stream, out, packed, suh, svh, k, n_out, packed_tiles_n, tile_n_offset, num_sms, err);
    else if (split == 2) reconstruct_had_tc_launch_s<K, cb, 2>(stream, out, packed, suh, svh, k, n_out, packed_tiles_n, tile_n_offset, num_sms, err);
    else reconstruct_had_tc_launch_s<K, cb, 4>(stream, out, packed, suh, svh, k, n_out, packed_tiles_n, tile_n_offset, num_sms, err);
}

// EXL3B_RECONSTRUCT_HAD=simt keeps the CUDA-core kernel (kernels_basic.cu), the twin the tests compare against;
// reconstruct_had_set_mode (test hook): 0 = environment / default, 1 = CUDA cores, 2 = tensor cores (2x = with x threads per row)
static int g_rt_mode = 0;
void reconstruct_had_set_mode(int mode)
{
    if (mode >= 20) { g_rt_mode = 2; g_rt_split = mode - 20; }         // 21 / 22 / 24: tensor cores with 1 / 2 / 4 threads per row
    else { g_rt_mode = mode; g_rt_split = 0; }
}
bool reconstruct_had_tc_enabled()
{
    if (g_rt_mode) return g_rt_mode == 2;
    static int on = -1;
    if (on < 0) { const char* e = getenv("EXL3B_RECONSTRUCT_HAD"); on = (e && e[0] == 's') ? 0 : 1; }
    return on != 0;
}

int launch_reconstruct_had_tc(cudaStream_t stream, half* unpacked, const uint16_t* packed, const half* suh, const half* svh,
                              int k, int n_out, int packed_tiles_n, int K, int cb, int64_t n_offset, int num_sms)
{
    if (k == 0 || n_out == 0) return 0;
    cudaError_t err = cudaSuccess;
    EXL3B_DISPATCH_K_CB(reconstruct_had_tc_launch, K, cb, stream, unpacked, packed, suh, svh, k, n_out, packed_tiles_n,
                        (int) (n_offset / 16), num_sms, &err);
    count_launch();
    EXL3B_CUDA(err);
    EXL3B_CUDA(cudaPeekAtLastError());
    return 0;
}

}  // namespace exl3b
