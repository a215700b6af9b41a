// Bit-faithful restatement of the single-precision libm routines the reference hot path calls:
//   cosf/sinf  (reference call sites: Cpp/Source/Maths.cpp:44-45, Cpp/Source/Test.cpp:116)
//   powf       (reference call site:  Cpp/Source/Maths.h:331, always powf(1-cosine, 5))
// The reference links the platform libm; on the parity platform that is glibc 2.39 (Ubuntu 2.39-0ubuntu8.5,
// x86-64), which is NOT part of /root/reference. glibc is not correctly rounded, and one
// differing ulp re-times a whole row's RNG stream, so the exact ("replay") mode needs the
// very same function. What follows restates glibc's published algorithms
//   sysdeps/ieee754/flt-32/{s_sinf.c, s_cosf.c, sincosf.h, sincosf_data.c}   (sinf/cosf, |x| < 120 path)
//   sysdeps/ieee754/flt-32/{e_powf.c, e_powf_log2_data.c, e_exp2f_data.c}    (powf)
// in the exact operation order of the x86-64 `*_fma` ifunc variants (the ones selected on any CPU with
// AVX2+FMA, i.e. every B200 host): those are compiled with FP contraction, so every `a*b + c` below that
// is a fused multiply-add in that build is an explicit fma() here, and every other operation is an
// explicitly rounded double op (never contracted by nvcc: __dmul_rn/__dadd_rn).
// Pinned by tests/test_libm.py: exhaustive equality with the platform libm over all 2^24 arguments the
// path can produce for sinf/cosf, and over all 2^32 floats for powf(x, 5).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDA_ARCH__)
#define TPT_HD __host__ __device__ __forceinline__
#define TPT_DMUL(a, b) __dmul_rn((a), (b))
#define TPT_DADD(a, b) __dadd_rn((a), (b))
#define TPT_DSUB(a, b) __dsub_rn((a), (b))
#define TPT_DFMA(a, b, c) __fma_rn((a), (b), (c))
#elif defined(__CUDACC__)
#define TPT_HD __host__ __device__ __forceinline__
#define TPT_DMUL(a, b) ((a) * (b))
#define TPT_DADD(a, b) ((a) + (b))
#define TPT_DSUB(a, b) ((a) - (b))
#define TPT_DFMA(a, b, c) fma((a), (b), (c))
#else
// Host-only build (tests/host_sim): compile with -ffp-contract=off so the plain ops stay separate.
#define TPT_HD inline
#define TPT_DMUL(a, b) ((a) * (b))
#define TPT_DADD(a, b) ((a) + (b))
#define TPT_DSUB(a, b) ((a) - (b))
#define TPT_DFMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

namespace tptlibm {

TPT_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
TPT_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
TPT_HD uint64_t d2u(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    union { double d; uint64_t u; } c; c.d = d; return c.u;
#endif
}
TPT_HD double u2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; uint64_t u; } c; c.u = u; return c.d;
#endif
}

// ---- sinf / cosf ----------------------------------------------------------------------------------
// glibc sincosf_data.c: __sincosf_table[2] = { sign[4], hpi_inv (2/pi * 2^24), hpi, c0..c4, s1..s3 };
// table[1] is table[0] with c0..c4/s1.. negated where the quadrant flips the sign.
// Polynomial coefficients (table 0). Table 1 differs by sign only: C0,C1,C2,C3,C4 -> -C0,-C1,-C2,-C3,-C4
// and S1..S3 unchanged.
#define TPT_SC_HPI_INV 0x1.45F306DC9C883p+23
#define TPT_SC_HPI 0x1.921FB54442D18p0
#define TPT_SC_C0 0x1p0
#define TPT_SC_C1 -0x1.ffffffd0c621cp-2
#define TPT_SC_C2 0x1.55553e1068f19p-5
#define TPT_SC_C3 -0x1.6c087e89a359dp-10
#define TPT_SC_C4 0x1.99343027bf8c3p-16
#define TPT_SC_S1 -0x1.555545995a603p-3
#define TPT_SC_S2 0x1.1107605230bc4p-7
#define TPT_SC_S3 -0x1.994eb3774cf24p-13

// sincosf.h sinf_poly, n even: sine polynomial. Argument already carries the quadrant sign.
TPT_HD float sin_poly(double x, double x2)
{
    double x3 = TPT_DMUL(x, x2);
    double s1 = TPT_DFMA(x2, TPT_SC_S3, TPT_SC_S2);   // p->s2 + x2 * p->s3
    double x7 = TPT_DMUL(x3, x2);
    double s = TPT_DFMA(x3, TPT_SC_S1, x);            // x + x3 * p->s1
    return (float)TPT_DFMA(x7, s1, s);                // s + x7 * s1
}
// sincosf.h sinf_poly, n odd: cosine polynomial; `flip` selects table[1] (all c negated).
// Table 1 negates every c coefficient; round-to-nearest is sign-symmetric, so every intermediate and the
// final rounding are the exact negation of the table-0 evaluation: evaluate once, flip the sign bit.
TPT_HD float cos_poly(double x2, bool flip)
{
    double x4 = TPT_DMUL(x2, x2);
    double c2 = TPT_DFMA(x2, TPT_SC_C4, TPT_SC_C3);   // p->c3 + x2 * p->c4
    double c1 = TPT_DFMA(x2, TPT_SC_C1, TPT_SC_C0);   // p->c0 + x2 * p->c1
    double x6 = TPT_DMUL(x4, x2);
    double c = TPT_DFMA(x4, TPT_SC_C2, c1);           // c1 + x4 * p->c2
    float r = (float)TPT_DFMA(x6, c2, c);             // c + x6 * c2
    return flip ? -r : r;
}

// sincosf.h reduce_fast (non-TOINT_INTRINSICS form, as built on x86-64)
TPT_HD double reduce_fast(double x, int* np)
{
    double r = TPT_DMUL(x, TPT_SC_HPI_INV);
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return TPT_DFMA(-(double)n, TPT_SC_HPI, x);                 // x - n * p->hpi (vfnmadd)
}

TPT_HD uint32_t abstop12(float x) { return (f2u(x) >> 20) & 0x7ff; }

// Returns false when |y| >= 120 or y is inf/nan (glibc's reduce_large / invalid path): the reference hot
// path never produces such arguments (they are 2*pi*[0,1)); callers fall back to the CUDA libm there.
TPT_HD bool sinf_glibc(float y, float* out)
{
    double x = (double)y;
    uint32_t top = abstop12(y);
    if (top < 0x3f4)                      // |y| < pi/4
    {
        if (top < 0x398) { *out = y; return true; }    // |y| < 2^-12
        double s = TPT_DMUL(x, x);
        *out = sin_poly(x, s);
        return true;
    }
    if (top < 0x42f)                      // |y| < 120
    {
        int n;
        x = reduce_fast(x, &n);
        double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[] = {1,-1,-1,1}
        bool flip = (n & 2) != 0;
        double x2 = TPT_DMUL(x, x);
        if ((n & 1) == 0) *out = sin_poly(TPT_DMUL(x, sgn), x2);
        else *out = cos_poly(x2, flip);
        return true;
    }
    return false;
}

TPT_HD bool cosf_glibc(float y, float* out)
{
    double x = (double)y;
    uint32_t top = abstop12(y);
    if (top < 0x3f4)
    {
        if (top < 0x398) { *out = 1.0f; return true; }
        double x2 = TPT_DMUL(x, x);
        *out = cos_poly(x2, false);
        return true;
    }
    if (top < 0x42f)
    {
        int n;
        x = reduce_fast(x, &n);
        double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
        bool flip = (n & 2) != 0;
        double x2 = TPT_DMUL(x, x);
        // cosf: sinf_poly(x * s, x * x, p, n ^ 1)
        if ((n & 1) != 0) *out = sin_poly(TPT_DMUL(x, sgn), x2);
        else *out = cos_poly(x2, flip);
        return true;
    }
    return false;
}

// ---- powf -------------------------------------------------------------------------------------------
// e_powf_log2_data.c: POWF_LOG2_TABLE_BITS = 4, POWF_LOG2_POLY_ORDER = 5, POWF_SCALE = 1 (x86-64 build)
// e_exp2f_data.c: EXP2F_TABLE_BITS = 5
} // namespace tptlibm

#if defined(__CUDACC__)
__device__ __constant__ const double tpt_d_log2_invc[16] = {
    0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0,
    0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
    0x1.0953f419900a7p+0, 0x1p+0, 0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
    0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
__device__ __constant__ const double tpt_d_log2_logc[16] = {
    -0x1.efec65b963019p-2, -0x1.b0b6832d4fca4p-2, -0x1.7418b0a1fb77bp-2, -0x1.39de91a6dcf7bp-2,
    -0x1.01d9bf3f2b631p-2, -0x1.97c1d1b3b7afp-3, -0x1.2f9e393af3c9fp-3, -0x1.960cbbf788d5cp-4,
    -0x1.a6f9db6475fcep-5, 0x0p+0, 0x1.338ca9f24f53dp-4, 0x1.476a9543891bap-3,
    0x1.e840b4ac4e4d2p-3, 0x1.40645f0c6651cp-2, 0x1.88e9c2c1b9ff8p-2, 0x1.ce0a44eb17bccp-2};
__device__ __constant__ const unsigned long long tpt_d_exp2_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
#endif

namespace tptlibm {

static const double h_log2_invc[16] = {
    0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0,
    0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
    0x1.0953f419900a7p+0, 0x1p+0, 0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
    0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
static const double h_log2_logc[16] = {
    -0x1.efec65b963019p-2, -0x1.b0b6832d4fca4p-2, -0x1.7418b0a1fb77bp-2, -0x1.39de91a6dcf7bp-2,
    -0x1.01d9bf3f2b631p-2, -0x1.97c1d1b3b7afp-3, -0x1.2f9e393af3c9fp-3, -0x1.960cbbf788d5cp-4,
    -0x1.a6f9db6475fcep-5, 0x0p+0, 0x1.338ca9f24f53dp-4, 0x1.476a9543891bap-3,
    0x1.e840b4ac4e4d2p-3, 0x1.40645f0c6651cp-2, 0x1.88e9c2c1b9ff8p-2, 0x1.ce0a44eb17bccp-2};
static const unsigned long long h_exp2_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

TPT_HD double tab_invc(int i)
{
#if defined(__CUDA_ARCH__)
    return tpt_d_log2_invc[i];
#else
    return h_log2_invc[i];
#endif
}
TPT_HD double tab_logc(int i)
{
#if defined(__CUDA_ARCH__)
    return tpt_d_log2_logc[i];
#else
    return h_log2_logc[i];
#endif
}
TPT_HD uint64_t tab_exp2(int i)
{
#if defined(__CUDA_ARCH__)
    return tpt_d_exp2_tab[i];
#else
    return h_exp2_tab[i];
#endif
}

// e_powf.c log2_inline, OFF = 0x3f330000, poly A[0..4]
TPT_HD double log2_inline(uint32_t ix)
{
    const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
                 A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
    uint32_t tmp = ix - 0x3f330000u;
    int i = (tmp >> (23 - 4)) % 16;
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)top >> 23;
    double invc = tab_invc(i);
    double logc = tab_logc(i);
    double z = (double)u2f(iz);
    double r = TPT_DFMA(z, invc, -1.0);          // z * invc - 1
    double y0 = TPT_DADD(logc, (double)k);
    double r2 = TPT_DMUL(r, r);
    double y = TPT_DFMA(A0, r, A1);
    double p = TPT_DFMA(A2, r, A3);
    double r4 = TPT_DMUL(r2, r2);
    double q = TPT_DFMA(A4, r, y0);
    q = TPT_DFMA(p, r2, q);
    y = TPT_DFMA(y, r4, q);
    return y;
}

// e_powf.c exp2_inline (SHIFT form), C = __exp2f_data.poly
TPT_HD float exp2_inline(double xd, uint32_t sign_bias)
{
    const double SHIFT = 0x1.8p+47;              // 0x1.8p+52 / N, N = 32
    const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    double kd = TPT_DADD(xd, SHIFT);
    uint64_t ki = d2u(kd);
    kd = TPT_DSUB(kd, SHIFT);
    double r = TPT_DSUB(xd, kd);
    uint64_t t = tab_exp2((int)(ki % 32));
    uint64_t ski = ki + sign_bias;
    t += ski << (52 - 5);
    double s = u2d(t);
    double z = TPT_DFMA(C0, r, C1);
    double r2 = TPT_DMUL(r, r);
    double y = TPT_DFMA(C2, r, 1.0);
    y = TPT_DFMA(z, r2, y);
    y = TPT_DMUL(y, s);
    return (float)y;
}

// e_powf.c checkint: 0 = not integer, 1 = odd integer, 2 = even integer
TPT_HD int checkint(uint32_t iy)
{
    int e = iy >> 23 & 0xff;
    if (e < 0x7f) return 0;
    if (e > 0x7f + 23) return 2;
    if (iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
    if (iy & (1u << (0x7f + 23 - e))) return 1;
    return 2;
}
TPT_HD bool zeroinfnan(uint32_t ix) { return 2 * ix - 1 >= 2u * 0x7f800000u - 1; }

// e_powf.c __powf, round-to-nearest mode, errno side effects dropped.
TPT_HD float powf_glibc(float x, float y)
{
    uint32_t sign_bias = 0;
    uint32_t ix = f2u(x), iy = f2u(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || zeroinfnan(iy))
    {
        if (zeroinfnan(iy))
        {
            if (2 * iy == 0) return 1.0f;                                     // issignaling ignored
            if (ix == 0x3f800000u) return 1.0f;
            if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return x + y;
            if (2 * ix == 2 * 0x3f800000u) return 1.0f;
            if ((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f; // |x|<1 && y==inf or |x|>1 && y==-inf
            return y * y;
        }
        if (zeroinfnan(ix))
        {
            float x2 = x * x;
            if ((ix & 0x80000000u) && checkint(iy) == 1) { x2 = -x2; sign_bias = 1; }
            if (2 * ix == 0 && (iy & 0x80000000u)) return sign_bias ? -INFINITY : INFINITY;
            return (iy & 0x80000000u) ? 1 / x2 : x2;
        }
        if (ix & 0x80000000u)
        {
            int yint = checkint(iy);
            if (yint == 0) return NAN;           // __math_invalidf
            if (yint == 1) sign_bias = 1u << (5 + 11);
            ix &= 0x7fffffffu;
        }
        if (ix < 0x00800000u)
        {
            ix = f2u(x * 0x1p23f);
            ix &= 0x7fffffffu;
            ix -= 23u << 23;
        }
    }
    double logx = log2_inline(ix);
    double ylogx = TPT_DMUL((double)y, logx);
    if ((d2u(ylogx) >> 47 & 0xffff) >= (d2u(126.0) >> 47))
    {
        if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? -INFINITY : INFINITY;      // __math_oflowf
        if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;                           // __math_uflowf
        if (ylogx < -149.0) return sign_bias ? -0x1p-149f : 0x1p-149f;                  // __math_may_uflowf
    }
    return exp2_inline(ylogx, sign_bias);
}

} // namespace tptlibm
