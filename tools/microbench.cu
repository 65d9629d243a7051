// Pipe-throughput microbenchmark for the decode budget (DESIGN.md "instruction budget").  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

enum { T_IMAD, T_LOP3, T_SHF, T_PRMT, T_DP4A, T_HFMA2, T_FFMA, T_IADD3, T_MIX_IMAD_LOP3, T_MIX_IMAD_DP4A,
       T_MIX_DECODE_CB2, T_MIX_DECODE_I8, T_SHFL, T_LDS, T_MIX_PRMT_IMAD, T_MIX_SHF_IMAD_DP4A, T_COUNT };
static const char* names[] = { "imad", "lop3", "shf", "prmt", "dp4a", "hfma2", "ffma", "iadd3", "imad+lop3",
    "imad+dp4a", "decode_cb2(shf,lop,imad,dp4a,.5prmt,.5hfma2)", "decode_i8(1.25alu+imad)", "shfl", "lds32",
    "prmt+imad", "shf+imad+dp4a" };
static const float ops_per_iter[] = { 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 5, 2.25f, 1, 1, 2, 3 };

template <int T>
__global__ void __launch_bounds__(1024) k(uint32_t* out, long long* cyc, uint32_t seed)
{
    __shared__ uint32_t sm[1024];
    uint32_t x[CHAINS];
    uint32_t y = seed ^ threadIdx.x, z = seed * 3 + 1;
    #pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 7 + c;
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < ITERS; ++i)
    {
        #pragma unroll
        for (int c = 0; c < CHAINS; ++c)
        {
            if (T == T_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            if (T == T_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[c]) : "r"(y), "r"(z));
            if (T == T_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y));
            if (T == T_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(x[c]) : "r"(y));
            if (T == T_DP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            if (T == T_HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            if (T == T_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*) &x[c]) : "f"(*(float*) &y), "f"(*(float*) &z));
            if (T == T_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(y));
            if (T == T_MIX_IMAD_LOP3)
            {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[c]) : "r"(y), "r"(z));
            }
            if (T == T_MIX_IMAD_DP4A)
            {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            }
            if (T == T_MIX_PRMT_IMAD)
            {
                asm volatile("prmt.b32 %0, %0, %1, 0x4410;" : "+r"(x[c]) : "r"(y));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            }
            if (T == T_MIX_SHF_IMAD_DP4A)
            {
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[c]) : "r"(y));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(z));
            }
            if (T == T_MIX_DECODE_CB2)
            {   // per weight: shf + lop(mask) + imad + dp4a ; per 2 weights: prmt + hfma2  (two weights per chain step)
                uint32_t a = x[c], b;
                asm volatile("shf.r.wrap.b32 %0, %1, %2, 7;" : "=r"(b) : "r"(a), "r"(y));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xc0;" : "+r"(b) : "r"(0xffffu), "r"(z));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(y), "r"(z));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(b) : "r"(0x01010101u), "r"(0x6400u));
                asm volatile("shf.r.wrap.b32 %0, %0, %1, 11;" : "+r"(a) : "r"(y));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xc0;" : "+r"(a) : "r"(0xffffu), "r"(z));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(y), "r"(z));
                asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(0x01010101u), "r"(0x6400u));
                asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a) : "r"(b));
                asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(0x1eee1eeeu), "r"(0xc931c931u));
                x[c] = a;
            }
            if (T == T_MIX_DECODE_I8)
            {   // per 4 weights: 1 shf + 4 prmt + 4 imad  => 2.25 ops per weight; here 4 weights per chain step
                uint32_t a = x[c], t, s0, s1, s2, s3;
                asm volatile("shf.r.wrap.b32 %0, %1, %2, 4;" : "=r"(t) : "r"(a), "r"(y));
                asm volatile("prmt.b32 %0, %1, %2, 0x4410;" : "=r"(s0) : "r"(a), "r"(0));
                asm volatile("prmt.b32 %0, %1, %2, 0x4421;" : "=r"(s1) : "r"(a), "r"(0));
                asm volatile("prmt.b32 %0, %1, %2, 0x4410;" : "=r"(s2) : "r"(t), "r"(0));
                asm volatile("prmt.b32 %0, %1, %2, 0x4421;" : "=r"(s3) : "r"(t), "r"(0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s0) : "r"(y), "r"(z));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s1) : "r"(y), "r"(s0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s2) : "r"(y), "r"(s1));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(s3) : "r"(y), "r"(s2));
                x[c] = s3;
            }
            if (T == T_SHFL) x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1);
            if (T == T_LDS) x[c] = sm[(x[c] + c) & 1023];
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    #pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T>
void run(uint32_t* out, long long* cyc, int sms, int threads)
{
    k<T><<<sms, threads>>>(out, cyc, 12345u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<T><<<sms, threads>>>(out, cyc, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    int mult = (T == T_MIX_DECODE_CB2) ? 2 : (T == T_MIX_DECODE_I8 ? 4 : 1);
    double units = (double) threads * ITERS * CHAINS * mult;            // "weights" or single ops per SM
    double inst = units * ops_per_iter[T] / 32.0;                       // warp instructions per SM
    printf("%-52s thr=%4d  %8.1f cyc  warp-inst/clk/SM %6.3f  units/clk/SM %7.2f  (%.3f ms, %.0f MHz)\n",
           names[T], threads, avg, inst / avg, units / avg, ms, avg / ms / 1e3);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * 256);
    for (int threads : {256, 512, 1024})
    {
        run<T_IMAD>(out, cyc, sms, threads); run<T_LOP3>(out, cyc, sms, threads); run<T_SHF>(out, cyc, sms, threads);
        run<T_PRMT>(out, cyc, sms, threads); run<T_DP4A>(out, cyc, sms, threads); run<T_HFMA2>(out, cyc, sms, threads);
        run<T_FFMA>(out, cyc, sms, threads); run<T_IADD3>(out, cyc, sms, threads);
        run<T_MIX_IMAD_LOP3>(out, cyc, sms, threads); run<T_MIX_IMAD_DP4A>(out, cyc, sms, threads);
        run<T_MIX_PRMT_IMAD>(out, cyc, sms, threads); run<T_MIX_SHF_IMAD_DP4A>(out, cyc, sms, threads);
        run<T_MIX_DECODE_CB2>(out, cyc, sms, threads); run<T_MIX_DECODE_I8>(out, cyc, sms, threads);
        run<T_SHFL>(out, cyc, sms, threads); run<T_LDS>(out, cyc, sms, threads);
        printf("\n");
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
