/*
 * CPU oracle, C restatement  --  TEST INFRASTRUCTURE, NOT PRODUCT.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may call this.
 *
 * Restates the reference's EXL3 decode (turboderp-org/exllamav3 @ 4f8ad012, paths under exllamav3/):
 *   window:    exllamav3_ext/quant/exl3_dq.cuh:15-31, scalar CPU form exllamav3_ext/cpu/moe_mul1.cpp:162-172
 *   codebooks: exllamav3_ext/quant/codebook.cuh:56-90
 *   tile perm: modules/quant/exl3_lib/quantize.py:22-44
 *   placement: exllamav3_ext/quant/reconstruct.cu:23-83
 * Pinned bit-exactly against oracle/exl3_oracle.py (tests/test_oracle.py::test_c_oracle_matches_numpy), which is
 * itself pinned against the reference's CUDA kernels (tests/golden/ref_gpu.npz).
 *
 * Built by oracle/Makefile into oracle/libexl3oracle.so (pthreads: used as the multi-threaded CPU baseline).
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

static float half_bits_to_float(uint16_t h)
{
    uint32_t s = (uint32_t) (h >> 15) << 31, e = (h >> 10) & 31, m = h & 1023, u;
    if (e == 0)
    {
        if (m == 0) u = s;
        else { e = 113; while (!(m & 1024)) { m <<= 1; e--; } u = s | (e << 23) | ((m & 1023) << 13); }
    }
    else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}

/* round-to-nearest-even double -> fp16 bits (values here are finite and well inside fp16 range) */
static uint16_t double_to_half_bits(double d)
{
    uint16_t sign = 0;
    if (d < 0 || (d == 0 && 1.0 / d < 0)) { sign = 0x8000; d = -d; }
    if (d == 0) return sign;
    int e; double fr = frexp(d, &e);            /* d = fr * 2^e, fr in [0.5, 1) */
    int exp = e - 1;                            /* d = (2 fr) * 2^exp */
    if (exp < -14)
    {
        double q = d * 16777216.0;              /* units of 2^-24 */
        double r = nearbyint(q);
        return sign | (uint16_t) r;             /* may carry into the normal range correctly */
    }
    double mant = fr * 2048.0;                  /* in [1024, 2048) */
    double r = nearbyint(mant);
    if (r >= 2048.0) { r = 1024.0; exp++; }
    if (exp > 15) return sign | 0x7c00;
    return sign | (uint16_t) (((exp + 15) << 10) | ((uint32_t) r - 1024));
}

static uint16_t decode_value(uint32_t state, int cb)
{
    uint32_t x;
    if (cb == 0) x = state * 89226354u + 64248484u;
    else if (cb == 1) x = state * 0xCBAC1FEDu;
    else x = state * 0x83DCD12Du;
    if (cb < 2)
    {
        x = (x & 0x8fff8fffu) ^ 0x3b603b60u;
        double s = (double) half_bits_to_float((uint16_t) (x & 0xffff)) + (double) half_bits_to_float((uint16_t) (x >> 16));
        return double_to_half_bits(s);
    }
    uint32_t sum = (x & 0xff) + ((x >> 8) & 0xff) + ((x >> 16) & 0xff) + (x >> 24);
    double h = (double) half_bits_to_float((uint16_t) (0x6400 + sum));
    double k_inv = (double) half_bits_to_float(0x1eee), k_bias = (double) half_bits_to_float(0xc931);
    return double_to_half_bits(h * k_inv + k_bias);
}

static void make_perm(int* perm)
{
    for (int t = 0; t < 32; ++t)
    {
        int r0 = (t % 4) * 2, r[4] = { r0, r0 + 1, r0 + 8, r0 + 9 }, c0 = t / 4;
        for (int i = 0; i < 8; ++i) perm[t * 8 + i] = r[i % 4] * 16 + (i < 4 ? c0 : c0 + 8);
    }
}

#include <pthread.h>

typedef struct { void* out; const uint16_t* trellis; int k, n, K, cb, f32, kt0, kt1; } job_t;

static void* run_job(void* arg)
{
    job_t* j = (job_t*) arg;
    int perm[256];
    make_perm(perm);
    const int tn = j->n / 16, nw = 8 * j->K, K = j->K, n = j->n;
    for (int kt = j->kt0; kt < j->kt1; ++kt)
        for (int nt = 0; nt < tn; ++nt)
        {
            const uint16_t* p = j->trellis + ((size_t) kt * tn + nt) * 16 * K;
            uint32_t words[64];
            memcpy(words, p, (size_t) 32 * K);
            for (int t = 0; t < 256; ++t)
            {
                int b0 = t * K + K - 16 + 256 * K, b1 = b0 + 16;
                int i0 = (b0 / 32) % nw, i1 = ((b1 - 1) / 32) % nw;
                int shift = ((b1 - 1) / 32 + 1) * 32 - b1;
                uint64_t merged = ((uint64_t) words[i0] << 32) | words[i1];
                uint32_t st = (uint32_t) (merged >> shift) & 0xffff;
                int r = perm[t] / 16, c = perm[t] % 16;
                size_t o = ((size_t) kt * 16 + r) * n + nt * 16 + c;
                uint16_t v = decode_value(st, j->cb);
                if (j->f32) ((float*) j->out)[o] = half_bits_to_float(v);
                else ((uint16_t*) j->out)[o] = v;
            }
        }
    return 0;
}

static void reconstruct_mt(void* out, const uint16_t* trellis, int k, int n, int K, int cb, int threads, int f32)
{
    const int tk = k / 16;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > tk) threads = tk > 0 ? tk : 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int i = 0; i < threads; ++i)
    {
        job_t j = { out, trellis, k, n, K, cb, f32, (int) ((long) tk * i / threads), (int) ((long) tk * (i + 1) / threads) };
        jobs[i] = j;
        if (threads == 1) run_job(&jobs[i]);
        else pthread_create(&th[i], 0, run_job, &jobs[i]);
    }
    if (threads > 1) for (int i = 0; i < threads; ++i) pthread_join(th[i], 0);
}

/* trellis (k/16, n/16, 16K) uint16 -> out (k, n) fp16 bits */
void exl3o_reconstruct(uint16_t* out, const uint16_t* trellis, int k, int n, int K, int cb, int threads)
{
    reconstruct_mt(out, trellis, k, n, K, cb, threads, 0);
}

/* same, fp32 output (the CPU baseline feeds a torch fp32 matmul like the reference's get_weight_tensor path) */
void exl3o_reconstruct_f32(float* out, const uint16_t* trellis, int k, int n, int K, int cb, int threads)
{
    reconstruct_mt(out, trellis, k, n, K, cb, threads, 1);
}
