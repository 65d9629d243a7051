// Scalar arithmetic of the int8 tensor-core codebook path (gemm_tc_i8_body.cuh), host + device: the activation -> digit
// decomposition its transform warps apply per value, and the reassembly its epilogue applies per accumulator pair.  Kept in
// one place so that the CPU tests can run exactly this code against the oracle's model of the path
// (tests/emu/decode_emu.cpp, tests/test_decode_emu.py); the kernels inline it (SASS unchanged by the extraction).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define EXL3B_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define EXL3B_HD inline
#endif

namespace exl3b {

constexpr int I8_QMAX = 32512;                                 // |q| <= 127 * 256 + 0  -> hi in [-127, 127]

EXL3B_HD int i8_rint(float x)
{
#ifdef __CUDA_ARCH__
    return __float2int_rn(x);
#else
    return (int) lrintf(x);                                    // round to nearest even (default rounding mode)
#endif
}

// One transformed activation value -> q = round(v * inv_scale) = 256 * hi + lo with BALANCED signed 8-bit digits
// (hi in [-127, 127], lo in [-128, 127]); each digit is replicated over the four product bytes of a weight, which is how
// the tensor core ends up summing them: sum_k digit_k * (b0 + b1 + b2 + b3)_k.
EXL3B_HD void i8_digits(float v, float inv_scale, int& q_sum, uint32_t& hi_w, uint32_t& lo_w)
{
    const int q = i8_rint(v * inv_scale);
    const int hi = (q + 128) >> 8;
    const int lo = q - (hi << 8);
    q_sum += q;                                                // the row's digit sum T = sum q_k goes into the epilogue
    hi_w = (uint32_t) (hi & 0xff) * 0x01010101u;
    lo_w = (uint32_t) (lo & 0xff) * 0x01010101u;
}

// One row's s32 accumulators of a column (d_hi = sum hi_k * bytesum_k, d_lo = sum lo_k * bytesum_k) and the row's digit sum
// T = sum q_k  ->  scale * sum_k q_k * w_k  with  w = k_inv * (1024 + bytesum) + k_bias, in two steps:
//     i8_centred_sum   sp = sum_k q_k (bytesum_k - 510), exact in 64-bit (the byte sum is centred to keep the fp32 part small)
//     i8_assemble      scale * (k_inv * sp + c1 * T),  c1 = (1024 + 510) * k_inv + k_bias
EXL3B_HD long long i8_centred_sum(int d_hi, int d_lo, int T)
{
    return 256ll * d_hi + (long long) d_lo - 510ll * T;
}
EXL3B_HD float i8_assemble(long long sp, int T, float scale, float k_inv, float c1)
{
    return scale * (k_inv * (float) sp + c1 * (float) T);
}

}  // namespace exl3b
