// Standard normals for fp_linear! from two 32-bit Philox words each, in fp32 with explicitly rounded operations only
// (add, mul, fma, div, sqrt -- all IEEE-754 correctly rounded), so that the CUDA kernel and the CPU oracle (which includes this
// header) produce bit-identical values.  The reference draws randn(rng) (collision_fp.jl:164-170); the draws are standardised
// exactly (mean 0, variance 1 over the cell, :182-211) before use, so fp32 resolution of the raw normals is ample
// and the fp64 log / sincos that used to dominate the kernel (1030 thread-instructions per particle) are gone.
//
// Box-Muller: u1 = (k1 + 0.5) 2^-24 in (0, 1), u2 = k2 2^-24 in [0, 1) from the top 24 bits of two words;
//   r = sqrt(-2 ln u1), (n0, n1) = r (cos 2 pi u2, sin 2 pi u2).
#pragma once
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define MBN_FN __device__ __forceinline__
#define MBN_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define MBN_MUL(a, b) __fmul_rn((a), (b))
#define MBN_ADD(a, b) __fadd_rn((a), (b))
#define MBN_DIV(a, b) __fdiv_rn((a), (b))
#define MBN_SQRT(a) __fsqrt_rn((a))
#define MBN_BITS(f) __float_as_uint((f))
#define MBN_FLOAT(u) __uint_as_float((u))
#else
#include <cmath>
#include <cstring>
#define MBN_FN inline
#define MBN_FMA(a, b, c) fmaf((a), (b), (c))
#define MBN_MUL(a, b) ((a) * (b))
#define MBN_ADD(a, b) ((a) + (b))
#define MBN_DIV(a, b) ((a) / (b))
#define MBN_SQRT(a) sqrtf((a))
static inline uint32_t mbn_bits_(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float mbn_float_(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#define MBN_BITS(f) mbn_bits_((f))
#define MBN_FLOAT(u) mbn_float_((u))
#endif

// natural logarithm of x in (0, 1]: x = m 2^e with m in [sqrt(1/2), sqrt(2)), ln m = 2 atanh(t), t = (m - 1) / (m + 1), |t| <= 0.1716
MBN_FN float mbn_log(float x) {
    uint32_t b = MBN_BITS(x);
    int e = (int)(b >> 23) - 127;
    b = (b & 0x007FFFFFu) | 0x3F800000u;  // m in [1, 2)
    float m = MBN_FLOAT(b);
    if (m > 1.41421356f) { m = MBN_MUL(m, 0.5f); e += 1; }
    const float t = MBN_DIV(MBN_ADD(m, -1.0f), MBN_ADD(m, 1.0f));
    const float t2 = MBN_MUL(t, t);
    float p = MBN_FMA(t2, 0.11111111f, 0.14285715f);  // 1/9, 1/7
    p = MBN_FMA(t2, p, 0.2f);
    p = MBN_FMA(t2, p, 0.33333334f);
    p = MBN_FMA(t2, p, 1.0f);
    const float lnm = MBN_MUL(MBN_MUL(2.0f, t), p);
    return MBN_FMA((float)e, 0.69314718f, lnm);
}

// (cos, sin) of 2 pi u, u in [0, 1): quadrant k = round(4 u), remainder a = u - k / 4 in [-1/8, 1/8], polynomials on |2 pi a| <= pi / 4
MBN_FN void mbn_sincos2pi(float u, float* c, float* s) {
    const float u4 = MBN_MUL(u, 4.0f);
    const int k = (int)MBN_ADD(u4, 0.5f);           // u4 in [0, 4): truncation == floor
    const float a = MBN_MUL(MBN_ADD(u4, -(float)k), 0.25f);  // exact: u4 - k is representable, * 0.25 is a power of two
    const float x = MBN_MUL(a, 6.2831855f);
    const float x2 = MBN_MUL(x, x);
    float ps = MBN_FMA(x2, -1.9841270e-4f, 8.3333338e-3f);  // -1/5040, 1/120
    ps = MBN_FMA(x2, ps, -0.16666667f);
    ps = MBN_FMA(x2, ps, 1.0f);
    const float sn = MBN_MUL(x, ps);
    float pc = MBN_FMA(x2, 2.4801588e-5f, -1.3888889e-3f);  // 1/40320, -1/720
    pc = MBN_FMA(x2, pc, 4.1666668e-2f);
    pc = MBN_FMA(x2, pc, -0.5f);
    const float cs = MBN_FMA(x2, pc, 1.0f);
    switch (k & 3) {
        case 0: *c = cs; *s = sn; break;
        case 1: *c = -sn; *s = cs; break;
        case 2: *c = -cs; *s = -sn; break;
        default: *c = sn; *s = -cs; break;
    }
}

// two normals from the words (w0, w1): w0 -> radius, w1 -> angle
MBN_FN void mbn_box_muller(uint32_t w0, uint32_t w1, float* n0, float* n1) {
    const float u1 = MBN_MUL(MBN_ADD((float)(w0 >> 8), 0.5f), 5.9604645e-8f);  // (k + 0.5) 2^-24 in (0, 1); exact
    const float u2 = MBN_MUL((float)(w1 >> 8), 5.9604645e-8f);                  // k 2^-24 in [0, 1); exact
    const float r = MBN_SQRT(MBN_MUL(-2.0f, mbn_log(u1)));
    float c, s;
    mbn_sincos2pi(u2, &c, &s);
    *n0 = MBN_MUL(r, c);
    *n1 = MBN_MUL(r, s);
}
