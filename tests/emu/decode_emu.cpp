// Host build of the library's trellis-decode templates (exllamav3_b200/csrc/decode.cuh) for the CPU tests: the same
// compile-time bit-window extraction and codebook arithmetic the CUDA kernels run per thread, with the handful of PTX
// instructions restated in C++.  Test infrastructure only (tests/test_decode_emu.py builds it with g++); not linked into
// the library.
#include <stdint.h>
#include <string.h>
#include <math.h>

#define EXL3B_HOST_EMU
#define __device__
#define __forceinline__ inline

// fma.rn.f16x2 / add.rn.f16x2: exact product and sum in double (22 + 11 significant bits fit), ONE rounding to fp16
static inline _Float16 h_from_bits(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return h; }
static inline uint16_t h_bits(_Float16 h) { uint16_t b; memcpy(&b, &h, 2); return b; }
static inline uint16_t f16_fma(uint16_t a, uint16_t b, uint16_t c)
{
    const double r = (double) h_from_bits(a) * (double) h_from_bits(b) + (double) h_from_bits(c);
    return h_bits((_Float16) r);
}
static inline uint32_t exl3b_emu_f16x2_fma(uint32_t a, uint32_t b, uint32_t c)
{
    return (uint32_t) f16_fma(a & 0xffff, b & 0xffff, c & 0xffff) | ((uint32_t) f16_fma(a >> 16, b >> 16, c >> 16) << 16);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift)
{
    const uint64_t v = ((uint64_t) hi << 32) | lo;
    return (uint32_t) (v >> (shift & 31));
}
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)          // unsigned bytes
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}

#include "../../exllamav3_b200/csrc/decode.cuh"
#include "../../exllamav3_b200/csrc/i8_math.cuh"
#include <vector>

using namespace exl3b;

template <int K, int cb>
static void dec16(int half, const uint32_t* w, uint32_t* out)
{
    uint32_t ww[K + 1], o[8];
    for (int i = 0; i <= K; ++i) ww[i] = w[i];
    if (half) decode16<K, cb, 1>(ww, o); else decode16<K, cb, 0>(ww, o);
    for (int i = 0; i < 8; ++i) out[i] = o[i];
}
template <int K>
static void dec16_i8(int half, const uint32_t* w, uint32_t* out)
{
    uint32_t ww[K + 1], o[16];
    for (int i = 0; i <= K; ++i) ww[i] = w[i];
    if (half) decode16_i8<K, 1>(ww, o); else decode16_i8<K, 0>(ww, o);
    for (int i = 0; i < 16; ++i) out[i] = o[i];
}
template <int K>
static void chunk_words(const uint32_t* tile, int chunk, uint32_t* w)
{
    uint32_t ww[K + 1];
    load_chunk<K>(tile, chunk, ww);
    for (int i = 0; i <= K; ++i) w[i] = ww[i];
}

#define FOR_K(FN, ...) switch (K) { case 1: FN<1>(__VA_ARGS__); break; case 2: FN<2>(__VA_ARGS__); break; case 3: FN<3>(__VA_ARGS__); break; \
    case 4: FN<4>(__VA_ARGS__); break; case 5: FN<5>(__VA_ARGS__); break; case 6: FN<6>(__VA_ARGS__); break; \
    case 7: FN<7>(__VA_ARGS__); break; case 8: FN<8>(__VA_ARGS__); break; default: return -1; }
template <int K> static void dec16_cb(int cb, int half, const uint32_t* w, uint32_t* out)
{
    if (cb == 0) dec16<K, 0>(half, w, out); else if (cb == 1) dec16<K, 1>(half, w, out); else dec16<K, 2>(half, w, out);
}

// The int8 tensor-core path's arithmetic for ONE activation row, end to end, from the device headers: row maximum -> digits
// (i8_digits) -> per weight the product word of decode16_i8 whose four bytes the tensor core multiplies with the replicated
// signed digit and sums into s32 -> i8_centred_sum / i8_assemble.  xh: the k transformed activations (fp16 values as float);
// tiles: (k/16) x (n/16) tiles of 8K words; acc: the n column sums BEFORE the output Hadamard.  Mirrors gemm_tc_i8_body.cuh
// (prologue maxima, digit warps, epilogue) without its parallel decomposition.
template <int K>
static int i8_row(const uint32_t* tiles, const float* xh, int k, int n, float* acc)
{
    float mx = 0.f;
    for (int i = 0; i < k; ++i) mx = fmaxf(mx, fabsf(xh[i]));
    const float inv_scale = mx > 0.f ? (float) I8_QMAX / mx : 0.f;
    const float scale = mx / (float) I8_QMAX;
    std::vector<int> hi(k), lo(k);
    int T = 0;
    for (int i = 0; i < k; ++i)
    {
        uint32_t hw, lw;
        i8_digits(xh[i], inv_scale, T, hw, lw);
        if (hw != (hw & 0xff) * 0x01010101u || lw != (lw & 0xff) * 0x01010101u) return -2;      // replicated over 4 bytes
        hi[i] = (int8_t) (hw & 0xff); lo[i] = (int8_t) (lw & 0xff);
        if (hi[i] < -127 || hi[i] > 127) return -3;
    }
    _Float16 hk, hb; uint16_t kb = 0x1eee, bb = 0xc931;
    memcpy(&hk, &kb, 2); memcpy(&hb, &bb, 2);
    const float k_inv = (float) hk, k_bias = (float) hb, c1 = 1534.0f * k_inv + k_bias;
    const int tn = n / 16;
    std::vector<long long> dh(n, 0), dl(n, 0);
    for (int kt = 0; kt < k / 16; ++kt)
        for (int nt = 0; nt < tn; ++nt)
        {
            const uint32_t* tile = tiles + ((size_t) kt * tn + nt) * 8 * K;
            for (int chunk = 0; chunk < 8; ++chunk)
            {
                uint32_t w[K + 1];
                load_chunk<K>(tile, chunk, w);
                for (int half = 0; half < 2; ++half)
                {
                    uint32_t o[16];
                    if (half) decode16_i8<K, 1>(w, o); else decode16_i8<K, 0>(w, o);
                    const int col = 16 * nt + chunk + 8 * half;
                    for (int r = 0; r < 16; ++r)
                    {
                        const int bs = (int) (o[r] & 0xff) + (int) ((o[r] >> 8) & 0xff) + (int) ((o[r] >> 16) & 0xff) + (int) (o[r] >> 24);
                        dh[col] += (long long) hi[16 * kt + r] * bs;
                        dl[col] += (long long) lo[16 * kt + r] * bs;
                    }
                }
            }
        }
    for (int c = 0; c < n; ++c)
    {
        if (dh[c] > 2147483647ll || dh[c] < -2147483648ll || dl[c] > 2147483647ll || dl[c] < -2147483648ll) return -4;   // s32 accumulators
        acc[c] = i8_assemble(i8_centred_sum((int) dh[c], (int) dl[c], T), T, scale, k_inv, c1);
    }
    return 0;
}

extern "C" {

// the K+1 words thread (chunk) reads from a tile of 8K words: preceding word (cyclic) + the chunk
int emu_load_chunk(int K, const uint32_t* tile, int chunk, uint32_t* w) { FOR_K(chunk_words, tile, chunk, w); return 0; }
// one tile column x 16 k-rows as 8 packed fp16 pairs (k = 2j, 2j+1): exact path (gemm_tc.cu, reconstruct, SIMT)
int emu_decode16(int K, int cb, int half, const uint32_t* w, uint32_t* out8) { if (cb < 0 || cb > 2) return -1; FOR_K(dec16_cb, cb, half, w, out8); return 0; }
// the same column as 16 raw products state * 0x83DCD12D in k order: int8 tensor-core path (gemm_tc_i8_body.cuh)
int emu_decode16_i8(int K, int half, const uint32_t* w, uint32_t* out16) { FOR_K(dec16_i8, half, w, out16); return 0; }
// K = 4 with a run-time half (branch-free experiment)
int emu_decode16_i8_k4_rt(int half, const uint32_t* w, uint32_t* out16)
{
    uint32_t ww[5], o[16];
    for (int i = 0; i < 5; ++i) ww[i] = w[i];
    decode16_i8_k4_rt(ww, half ? 0u : 16u, o);
    for (int i = 0; i < 16; ++i) out16[i] = o[i];
    return 0;
}
int emu_i8_row(int K, const uint32_t* tiles, const float* xh, int k, int n, float* acc)
{
    switch (K) { case 1: return i8_row<1>(tiles, xh, k, n, acc); case 2: return i8_row<2>(tiles, xh, k, n, acc);
                 case 3: return i8_row<3>(tiles, xh, k, n, acc); case 4: return i8_row<4>(tiles, xh, k, n, acc);
                 case 5: return i8_row<5>(tiles, xh, k, n, acc); case 6: return i8_row<6>(tiles, xh, k, n, acc);
                 case 7: return i8_row<7>(tiles, xh, k, n, acc); case 8: return i8_row<8>(tiles, xh, k, n, acc); }
    return -1;
}

// thread -> column mapping inside a 128-column strip
int emu_strip_col(int q, int lane) { return strip_col(q, lane); }

}
