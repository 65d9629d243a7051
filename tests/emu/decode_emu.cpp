// Host build of the library's trellis-decode templates (exllamav3_b200/csrc/decode.cuh) for the CPU tests: the same
// compile-time bit-window extraction and codebook arithmetic the CUDA kernels run per thread, with the handful of PTX
// instructions restated in C++.  Test infrastructure only (tests/test_decode_emu.py builds it with g++); not linked into
// the library.
#include <stdint.h>
#include <string.h>

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
// thread -> column mapping inside a 128-column strip
int emu_strip_col(int q, int lane) { return strip_col(q, lane); }

}
