// Device math shared by the dmfg kernels (sm_100a).
//
//   * Philox4x32 counter-based generator (Salmon et al., SC'11; 7 rounds in the Gamma sampler, 10
//     elsewhere) -- counters are
//     (population id, slot, attempt) so a draw never depends on the GPU count,
//     the grid shape or the kernel variant.
//   * Box-Muller normals + Marsaglia-Tsang Gamma(shape,1) with the U^(1/a) boost
//     for shape < 1 -- the in-kernel replacement of np.random.gamma
//     (mfg_ac2.py:242); validated statistically, not bit-for-bit.
//   * softplus concentration alpha / d alpha / d theta (mfg_ac2.py:228-234)
//   * digamma (scipy.special.digamma at mfg_ac2.py:364-367)
//
// Tm is the "math type": float for the throughput path, double for the
// float64 parity path.  All REDUCTIONS are done in double by the callers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmfg {

// ------------------------------------------------------------------ Philox
#define DMFG_PHILOX_M0 0xD2511F53u
#define DMFG_PHILOX_M1 0xCD9E8D57u
#define DMFG_PHILOX_W0 0x9E3779B9u
#define DMFG_PHILOX_W1 0xBB67AE85u

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                      uint32_t& c3, uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(DMFG_PHILOX_M0, c0);
    const uint32_t hi1 = __umulhi(DMFG_PHILOX_M1, c2);
#else
    const uint32_t hi0 = (uint32_t)(((uint64_t)DMFG_PHILOX_M0 * c0) >> 32);
    const uint32_t hi1 = (uint32_t)(((uint64_t)DMFG_PHILOX_M1 * c2) >> 32);
#endif
    const uint32_t lo0 = DMFG_PHILOX_M0 * c0;
    const uint32_t lo1 = DMFG_PHILOX_M1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
}

template <int ROUNDS>
__host__ __device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                     uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += DMFG_PHILOX_W0;
        k1 += DMFG_PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}
// Philox4x32-10: start rows of the learners, dropout masks, host-side draws
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}
// The Gamma sampler -- 225 variates per population-step, a quarter of the step's instructions -- runs
// Philox4x32-7: the same round function and key schedule with 7 rounds, the smallest round count Salmon et
// al. (SC'11, table 2) report as passing the full BigCrush battery ("Crush-resistant"); 10 is their default
// with a safety margin.  Measured on B200: -8 % on the whole train step.
#ifndef DMFG_GAMMA_ROUNDS
#define DMFG_GAMMA_ROUNDS 7
#endif
__host__ __device__ __forceinline__ uint4 philox_gamma(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1) {
    return philox4x32<DMFG_GAMMA_ROUNDS>(c0, c1, c2, c3, k0, k1);
}

// The same generator with the round keys (k + r*W) precomputed: the kernels take them as launch
// parameters so the key schedule costs nothing per call; 64-bit products map to one IMAD.WIDE each.
struct PhiloxKeys { uint32_t k[2 * DMFG_GAMMA_ROUNDS]; };
__host__ __device__ __forceinline__ PhiloxKeys make_philox_keys(uint64_t seed) {
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < DMFG_GAMMA_ROUNDS; ++r) { K.k[2 * r] = k0; K.k[2 * r + 1] = k1; k0 += DMFG_PHILOX_W0; k1 += DMFG_PHILOX_W1; }
    return K;
}
__device__ __forceinline__ uint4 philox_gamma(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& K) {
#pragma unroll
    for (int r = 0; r < DMFG_GAMMA_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)DMFG_PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)DMFG_PHILOX_M1 * c2;
        c0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r];
        c1 = (uint32_t)p1;
        c2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
        c3 = (uint32_t)p0;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Identifies one stream of draws: key = seed, counter words 0/1 = global population id.
struct NoiseKey {
    uint32_t k0, k1;     // seed
    uint32_t p0, p1;     // population (or learner) id
};
__host__ __device__ __forceinline__ NoiseKey make_noise_key(uint64_t seed, uint64_t pop) {
    NoiseKey k;
    k.k0 = (uint32_t)seed; k.k1 = (uint32_t)(seed >> 32);
    k.p0 = (uint32_t)pop;  k.p1 = (uint32_t)(pop >> 32);
    return k;
}

// Slot of the Gamma pair (columns 2p, 2p+1) of row i at (global) step t.
__host__ __device__ __forceinline__ uint32_t gamma_slot(uint32_t t, int d, int i, int p) {
    const uint32_t pd = (uint32_t)((d + 1) >> 1);
    return (t * (uint32_t)d + (uint32_t)i) * pd + (uint32_t)p;
}
#define DMFG_CTR_BOOST 0x80000000u   // counter word 3 of the boost uniforms (attempts count from 0)
#define DMFG_CTR_START 0xC0000000u   // counter word 3 of the start-row draw of the learners

// ---- MUFU-level primitives (approx, flush-to-zero): 1 instruction each on the XU pipe ----------------
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sin_approx(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cos_approx(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#define DMFG_LN2 0.69314718055994531f
#define DMFG_LOG2E 1.4426950408889634f

// uniform in (0,1), never 0 or 1
__device__ __forceinline__ float u01(uint32_t w) {
    return fmaf(__uint2float_rz(w), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// 23-bit uniform in [0,1) without an int->float conversion (those run on the quarter-rate XU pipe):
// the top 23 bits become the mantissa of a float in [1,2).
__device__ __forceinline__ float u01_23(uint32_t w) { return __uint_as_float((w >> 9) | 0x3f800000u) - 1.0f; }

// two standard normals from two 32-bit words (Box-Muller).  The angle is taken in (-pi, pi), where
// sin/cos.approx are accurate to ~5e-7 absolute; the radius resolves 32 bits (tails to 6.6 sigma).
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& n0, float& n1) {
    const float r = sqrt_approx(-2.0f * DMFG_LN2 * lg2_approx(u01(w0)));
    // angle in (-pi, pi) from the top 23 bits of w1: [1,2) * 2pi - 3pi (+ half a step so that neither end is hit)
    const float ang = fmaf(__uint_as_float((w1 >> 9) | 0x3f800000u), 6.2831845f, -9.4247770f);
    n0 = r * cos_approx(ang);
    n1 = r * sin_approx(ang);
}

// Marsaglia-Tsang proposal for Gamma(dd + 1/3, 1), cc = 1/sqrt(9 dd): y = dd (1 + cc x)^3, accepted iff
// ln u < x^2/2 + dd (1 - v + ln v), v = (1 + cc x)^3.
// With e = cc x the right-hand side is  -dd (0.75 e^4 - 0.6 e^5 + 0.5 e^6 - ...)  (dd cc^2 = 1/9 cancels the
// x^2/2 term exactly), bounded below by -1.5 dd e^4 for |e| <= 1/2.  Since exp(-z) >= 1 - z,
//     |e| <= 1/2  and  u < thr = 1 - 1.5 dd e^4      ==> accept
// is a squeeze that fails with probability ~0.055/dd only (Marsaglia & Tsang's generic 0.0331 x^4 squeeze
// fails 10 % of the time for every shape).  __fmul_rn / __fadd_rn pin the roundings, so every kernel
// variant computes bit-identical proposals.
__device__ __forceinline__ bool mt_squeeze(float dd, float cc, float x, float u, float& y, float& thr) {
    const float e = __fmul_rn(cc, x);
    const float v1 = __fmaf_rn(cc, x, 1.0f);       // every mul-add site of the proposal is an EXPLICIT fma, so
                                                   // scalar and packed (f32x2) code cannot round differently
    const float e2 = __fmul_rn(e, e);
    y = __fmul_rn(dd, __fmul_rn(v1, __fmul_rn(v1, v1)));
    thr = __fmaf_rn(__fmul_rn(-1.5f, dd), __fmul_rn(e2, e2), 1.0f);
    return (fabsf(e) <= 0.5f) & (u < thr);
}
// the exact acceptance test (only reached when the squeeze failed), cancellation-free
__device__ __forceinline__ bool mt_exact(float dd, float cc, float x, float u) {
    const float e = __fmul_rn(cc, x);
    const float v1 = __fmaf_rn(cc, x, 1.0f);
    if (v1 <= 0.0f) return false;
    const float e2 = e * e;
    float rhs;
    if (fabsf(e) < 0.25f) {
        // sum_{k>=4} (-1)^(k+1) e^k/k = -e^4 h,  h = sum_{m=0..9} (-e)^m/(m+4)  (|e|<1/4: rel. err < 2e-7)
        float h = 1.0f / 13.0f;
        h = fmaf(h, -e, 1.0f / 12.0f);
        h = fmaf(h, -e, 1.0f / 11.0f);
        h = fmaf(h, -e, 1.0f / 10.0f);
        h = fmaf(h, -e, 1.0f / 9.0f);
        h = fmaf(h, -e, 1.0f / 8.0f);
        h = fmaf(h, -e, 1.0f / 7.0f);
        h = fmaf(h, -e, 1.0f / 6.0f);
        h = fmaf(h, -e, 1.0f / 5.0f);
        h = fmaf(h, -e, 1.0f / 4.0f);
        rhs = -3.0f * dd * e2 * e2 * h;
    } else {
        const float v = v1 * v1 * v1;
        rhs = 0.5f * x * x + dd * (1.0f - v + __logf(v));
    }
    return __logf(u) < rhs;
}

// Shapes below 1: Gamma(a) = Gamma(a+1) U^(1/a) with U uniform and independent of the Gamma(a+1) draw
// (a == 0 gives 0 like np.random.gamma).  When the proposal was accepted by the squeeze, u | accept is
// uniform on (0, thr) whatever x was, so U = u / thr costs nothing; otherwise a fresh word is drawn.
__device__ __forceinline__ float boost_apply(float y, float a, float U) {
    return a > 0.0f ? __fmul_rn(y, ex2_approx(__fmul_rn(lg2_approx(U), rcp_approx(a)))) : 0.0f;
}
__device__ __forceinline__ float squeeze_uniform(float u, float thr) { return __fmul_rn(u, rcp_approx(thr)); }

struct GammaSetup {
    float dd, cc;
    bool boost;
};
// shape = alpha * scale; shapes below 1 are sampled as Gamma(shape + 1) and boosted afterwards.
// dd = shape (+1) - 1/3 as ONE fma of (alpha, scale)
__device__ __forceinline__ GammaSetup gamma_setup(float alpha, float scale) {
    GammaSetup g;
    g.boost = __fmul_rn(alpha, scale) < 1.0f;
    g.dd = __fmaf_rn(alpha, scale, g.boost ? (2.0f / 3.0f) : -(1.0f / 3.0f));
    g.cc = rsqrt_approx(__fmul_rn(9.0f, g.dd));
    return g;
}

// Gamma(a0,1), Gamma(a1,1) for the pair in `slot` -- the reference form of the sampler (generic / parity
// kernels, and the out-of-line path of the throughput kernel).  One Philox call per attempt feeds both
// elements (two Box-Muller normals + two acceptance uniforms); rejected elements move on to attempt+1.
__device__ __forceinline__ void gamma_pair(const NoiseKey& nk, uint32_t slot, float al0, float al1, float scale,
                                           float& y0, float& y1) {
    const GammaSetup g0 = gamma_setup(al0, scale), g1 = gamma_setup(al1, scale);
    const float a0 = __fmul_rn(al0, scale), a1 = __fmul_rn(al1, scale);
    bool done0 = false, done1 = false, sq0 = false, sq1 = false;
    float ub0 = 0.5f, ub1 = 0.5f;
    y0 = 0.0f; y1 = 0.0f;
    uint32_t attempt = 0;
    do {
        const uint4 w = philox_gamma(nk.p0, nk.p1, slot, attempt, nk.k0, nk.k1);
        float n0, n1, thr;
        box_muller(w.x, w.y, n0, n1);
        if (!done0) {
            const float u = u01_23(w.z);
            if (mt_squeeze(g0.dd, g0.cc, n0, u, y0, thr)) { done0 = sq0 = true; ub0 = squeeze_uniform(u, thr); }
            else done0 = mt_exact(g0.dd, g0.cc, n0, u);
        }
        if (!done1) {
            const float u = u01_23(w.w);
            if (mt_squeeze(g1.dd, g1.cc, n1, u, y1, thr)) { done1 = sq1 = true; ub1 = squeeze_uniform(u, thr); }
            else done1 = mt_exact(g1.dd, g1.cc, n1, u);
        }
        ++attempt;
    } while (!(done0 && done1) && attempt < 64u);
    if ((g0.boost && !sq0) || (g1.boost && !sq1)) {
        const uint4 w = philox_gamma(nk.p0, nk.p1, slot, DMFG_CTR_BOOST, nk.k0, nk.k1);
        if (!sq0) ub0 = u01(w.x);
        if (!sq1) ub1 = u01(w.y);
    }
    if (g0.boost) y0 = boost_apply(y0, a0, ub0);
    if (g1.boost) y1 = boost_apply(y1, a1, ub1);
}

// ---- branch-light variant for the throughput kernel --------------------------------------------------
// Same draws as gamma_pair (same counters, roundings and accept decisions): the straight-line part covers
// attempt 0 accepted by the squeeze, including shapes below 1; a squeeze miss (~0.06/shape) re-runs the pair
// out of line.  One rarely taken branch per pair.
static __device__ __noinline__ float2 gamma_pair_redo(uint32_t p0, uint32_t p1, uint32_t k0, uint32_t k1,
                                                      uint32_t slot, float al0, float al1, float scale) {
    NoiseKey nk;
    nk.k0 = k0; nk.k1 = k1; nk.p0 = p0; nk.p1 = p1;
    float y0, y1;
    gamma_pair(nk, slot, al0, al1, scale, y0, y1);
    return make_float2(y0, y1);
}
// packed single precision (Blackwell FFMA2 / FMUL2 / FADD2: two lanes of math per issue slot)
#ifdef DMFG_AB_SCALAR_ALL   // A/B probe: every packed operation as two scalar ones (same roundings)
__device__ __forceinline__ float2 ab_ffma2(float2 a, float2 b, float2 c) { return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
__device__ __forceinline__ float2 ab_fmul2(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float2 ab_fadd2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
#define __ffma2_rn ab_ffma2
#define __fmul2_rn ab_fmul2
#define __fadd2_rn ab_fadd2
#endif
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// y == 0 -> 1e-20 (mfg_ac2.py:244); 0 < y < FLT_MIN (denormal) -> FLT_MIN
__device__ __forceinline__ float gamma_floor(float y) { return y == 0.0f ? 1e-20f : fmaxf(y, 1.17549435e-38f); }

__device__ __forceinline__ void gamma_pair_fast(const NoiseKey& nk, const PhiloxKeys& K, uint32_t slot, float2 al,
                                                float scale, float& y0, float& y1) {
    const uint4 w = philox_gamma(nk.p0, nk.p1, slot, 0u, K);
    // Box-Muller: (n0, n1) = r (cos, sin)
    const float r = sqrt_approx(-2.0f * DMFG_LN2 * lg2_approx(u01(w.x)));
    const float ang = fmaf(__uint_as_float((w.y >> 9) | 0x3f800000u), 6.2831845f, -9.4247770f);
    const float2 n = __fmul2_rn(splat2(r), make_float2(cos_approx(ang), sin_approx(ang)));
    // gamma_setup for both elements
    const float2 a = __fmul2_rn(al, splat2(scale));
    // shapes below 1 (sampled as shape + 1, boosted afterwards) are rare: one min + one test per pair decides
    // whether the offset of dd needs its per-element select
#ifdef DMFG_AB_NOSMALL
    const bool small = false;          // A/B probe only (wrong for shapes < 1): cost of the small-shape handling
#else
    const bool small = fminf(a.x, a.y) < 1.0f;
#endif
    float2 off = splat2(-(1.0f / 3.0f));
    if (small) off = make_float2(a.x < 1.0f ? (2.0f / 3.0f) : -(1.0f / 3.0f), a.y < 1.0f ? (2.0f / 3.0f) : -(1.0f / 3.0f));
    const float2 dd = __ffma2_rn(al, splat2(scale), off);
    const float2 dd9 = __fmul2_rn(splat2(9.0f), dd);
    const float2 cc = make_float2(rsqrt_approx(dd9.x), rsqrt_approx(dd9.y));
    // mt_squeeze for both elements
    const float2 e = __fmul2_rn(cc, n);
    const float2 v1 = __ffma2_rn(cc, n, splat2(1.0f));
    const float2 e2 = __fmul2_rn(e, e);
    const float2 y = __fmul2_rn(dd, __fmul2_rn(v1, __fmul2_rn(v1, v1)));
    const float2 thr = __ffma2_rn(__fmul2_rn(splat2(-1.5f), dd), __fmul2_rn(e2, e2), splat2(1.0f));
    // acceptance uniforms: top 23 bits as the mantissa of [1,2), minus 1 -- no int->float conversion (XU pipe)
    const float2 u = __fadd2_rn(make_float2(__uint_as_float((w.z >> 9) | 0x3f800000u),
                                            __uint_as_float((w.w >> 9) | 0x3f800000u)), splat2(-1.0f));
    const bool ok0 = (fabsf(e.x) <= 0.5f) & (u.x < thr.x);
    const bool ok1 = (fabsf(e.y) <= 0.5f) & (u.y < thr.y);
    y0 = y.x;
    y1 = y.y;
    // A squeeze-accepted variate without boost is dd v^3 with v >= 1/2: never 0.  Only the two rare paths can
    // underflow to 0, so the reference's y == 0 -> 1e-20 substitution (mfg_ac2.py:244) lives there.
    // A boosted variate can also land in the float denormal range (y U^(1/a) with a tiny shape): the MUFU units
    // flush denormals to zero (lg2 -> -inf), so those are lifted to the smallest normal float -- they stand for
    // transition probabilities below 1e-38 either way (found by the random-regime test at theta = 28).
    if (!(ok0 & ok1)) {
        const float2 yy = gamma_pair_redo(nk.p0, nk.p1, nk.k0, nk.k1, slot, al.x, al.y, scale);
        y0 = gamma_floor(yy.x);
        y1 = gamma_floor(yy.y);
    } else if (small) {
        if (a.x < 1.0f) y0 = gamma_floor(boost_apply(y0, a.x, squeeze_uniform(u.x, thr.x)));
        if (a.y < 1.0f) y1 = gamma_floor(boost_apply(y1, a.y, squeeze_uniform(u.y, thr.y)));
    }
}

// ------------------------------------------------------------ policy alpha
// x = pi_j - pi_i - shift;  alpha = ln(1+exp(theta x));  alpha' = x / (1+exp(-theta x))
// (mfg_ac2.py:228-234), evaluated in the overflow-free form.
template <typename Tm> struct Math;
template <> struct Math<float> {
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float log1p_(float x) { return log1pf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
};
template <> struct Math<double> {
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double log1p_(double x) { return log1p(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
};

template <typename Tm>
__device__ __forceinline__ void policy_alpha(Tm theta, Tm x, Tm& alpha, Tm& alpha_deriv) {
    const Tm t = theta * x;
    const Tm e = Math<Tm>::exp_(-Math<Tm>::abs_(t));       // in (0,1]
    const Tm l = Math<Tm>::log1p_(e);
    const Tm inv = Tm(1) / (Tm(1) + e);
    const bool pos = t >= Tm(0);
    alpha = pos ? t + l : l;
    alpha_deriv = x * (pos ? inv : e * inv);
}

// ----------------------------------------------------------------- digamma
// psi(x), x > 0.  Recurrence psi(x) = psi(x+n) - sum_{k<n} 1/(x+k), then the
// asymptotic series.  float: n = 6 folded into ONE division via
//   x(x+5) = q, (x+1)(x+4) = q+4, (x+2)(x+3) = q+6,  sum = q'(3q^2+20q+24) / (q(q+4)(q+6)),
// every term positive, so small x (psi ~ -1/x, alpha down to 1e-5) keeps full
// relative accuracy.
__device__ __forceinline__ float digamma(float x) {
    float s = 0.0f, z = x;
    if (x < 1.0e4f) {
        const float q = fmaf(x, x, 5.0f * x);
        const float num = fmaf(2.0f, x, 5.0f) * fmaf(fmaf(3.0f, q, 20.0f), q, 24.0f);
        const float den = q * (q + 4.0f) * (q + 6.0f);
        s = num / den;
        z = x + 6.0f;
    }
    const float rz = 1.0f / z;
    const float r2 = rz * rz;
    // 1/12 - r2/120 + r2^2/252 - r2^3/240
    float p = fmaf(r2, -1.0f / 240.0f, 1.0f / 252.0f);
    p = fmaf(r2, p, -1.0f / 120.0f);
    p = fmaf(r2, p, 1.0f / 12.0f);
    return logf(z) - 0.5f * rz - r2 * p - s;
}

__device__ __forceinline__ double digamma(double x) {
    double s = 0.0;
    while (x < 10.0) { s += 1.0 / x; x += 1.0; }
    const double rz = 1.0 / x;
    const double r2 = rz * rz;
    // Bernoulli series: 1/12, 1/120, 1/252, 1/240, 1/132, 691/32760, 1/12
    double p = 1.0 / 12.0;
    p = fma(r2, -p, 691.0 / 32760.0);
    p = fma(r2, -p, 1.0 / 132.0);
    p = fma(r2, -p, 1.0 / 240.0);
    p = fma(r2, -p, 1.0 / 252.0);
    p = fma(r2, -p, 1.0 / 120.0);
    p = fma(r2, -p, 1.0 / 12.0);
    return log(x) - 0.5 * rz - r2 * p - s;
}

// ln(P) with the reference's P == 0 -> 1e-100 substitution (mfg_ac2.py:369).
__device__ __forceinline__ float log_prob(float p) { return p > 0.0f ? logf(p) : -230.25850929940458f; }
__device__ __forceinline__ double log_prob(double p) { return p > 0.0 ? log(p) : -230.25850929940458; }

// log1p(e) / e on (0, 1]: degree-8 interpolant at the Chebyshev nodes (highest power first: C8 .. C0)
#define DMFG_L1P_C8 5.126102141e-03f
#define DMFG_L1P_C7 -2.907406468e-02f
#define DMFG_L1P_C6 7.751608674e-02f
#define DMFG_L1P_C5 -1.360224762e-01f
#define DMFG_L1P_C4 1.907688074e-01f
#define DMFG_L1P_C3 -2.483539899e-01f
#define DMFG_L1P_C2 3.331812171e-01f
#define DMFG_L1P_C1 -4.999944498e-01f
#define DMFG_L1P_C0 9.999999659e-01f
// ------------------------------------------------- fast float variants (throughput kernels)
// Same functions as policy_alpha<float> / digamma(float) built from single MUFU operations; relative
// error <= ~3e-6 on alpha, alpha', psi over the operating range (tests/test_device_math_gpu.py).
__device__ __forceinline__ void policy_alpha_fast(float theta, float x, float& alpha, float& alpha_deriv) {
    const float t = theta * x;
    const float e = ex2_approx(-fabsf(t) * DMFG_LOG2E);             // exp(-|t|) in (0,1]
    const float u = 1.0f + e;
    // log1p(e) = e * p8(e) on (0, 1]: degree-8 interpolant at the Chebyshev nodes, relative error 2.1e-7 in
    // float over the whole range (alpha down to exp(-20)) -- no MUFU, no branch between a series and lg2
    float pl = fmaf(DMFG_L1P_C8, e, DMFG_L1P_C7);
    pl = fmaf(pl, e, DMFG_L1P_C6);
    pl = fmaf(pl, e, DMFG_L1P_C5);
    pl = fmaf(pl, e, DMFG_L1P_C4);
    pl = fmaf(pl, e, DMFG_L1P_C3);
    pl = fmaf(pl, e, DMFG_L1P_C2);
    pl = fmaf(pl, e, DMFG_L1P_C1);
    pl = fmaf(pl, e, DMFG_L1P_C0);
    const float l = __fmul_rn(pl, e);
    const float ru = rcp_approx(u);
    const bool pos = t >= 0.0f;
    alpha = fmaxf(t, 0.0f) + l;
    alpha_deriv = x * ((pos ? 1.0f : e) * ru);
}

// psi(x), x > 0: 4-step recurrence folded into one division, x(x+3) = q, (x+1)(x+2) = q+2:
//   sum_{k<4} 1/(x+k) = (2x+3)(2q+2) / (q (q+2)),   then the asymptotic series at z = x + 4 (error < 7e-8).
// One reciprocal serves both the recurrence and 1/z.
__device__ __forceinline__ float digamma_fast(float x) {
    const float q = fmaxf(fmaf(x, x, 3.0f * x), 1e-30f);
    const float num = fmaf(2.0f, x, 3.0f) * fmaf(2.0f, q, 2.0f);
    const float den = q * (q + 2.0f);
    const float z = x + 4.0f;
    const float R = rcp_approx(den * z);
    const float s = num * (z * R);
    const float rz = den * R;
    const float r2 = rz * rz;
    float p = fmaf(r2, -1.0f / 252.0f, 1.0f / 120.0f);
    p = fmaf(r2, p, -1.0f / 12.0f);
    float res = fmaf(lg2_approx(z), DMFG_LN2, -s);
    res = fmaf(-0.5f, rz, res);
    return fmaf(r2, p, res);
}

// packed log1p(e), same fma chain as the scalar form (bit-identical per element)
__device__ __forceinline__ float2 log1p_poly2(float2 e) {
#ifdef DMFG_AB_SCALAR_L1P
    float2 o;                       // A/B probe: the same chain as two scalar FFMA streams (FMA-lite can take them)
    {
        float pl = fmaf(DMFG_L1P_C8, e.x, DMFG_L1P_C7);
        pl = fmaf(pl, e.x, DMFG_L1P_C6); pl = fmaf(pl, e.x, DMFG_L1P_C5); pl = fmaf(pl, e.x, DMFG_L1P_C4);
        pl = fmaf(pl, e.x, DMFG_L1P_C3); pl = fmaf(pl, e.x, DMFG_L1P_C2); pl = fmaf(pl, e.x, DMFG_L1P_C1);
        pl = fmaf(pl, e.x, DMFG_L1P_C0);
        o.x = __fmul_rn(pl, e.x);
    }
    {
        float pl = fmaf(DMFG_L1P_C8, e.y, DMFG_L1P_C7);
        pl = fmaf(pl, e.y, DMFG_L1P_C6); pl = fmaf(pl, e.y, DMFG_L1P_C5); pl = fmaf(pl, e.y, DMFG_L1P_C4);
        pl = fmaf(pl, e.y, DMFG_L1P_C3); pl = fmaf(pl, e.y, DMFG_L1P_C2); pl = fmaf(pl, e.y, DMFG_L1P_C1);
        pl = fmaf(pl, e.y, DMFG_L1P_C0);
        o.y = __fmul_rn(pl, e.y);
    }
    return o;
#endif
    float2 pl = __ffma2_rn(splat2(DMFG_L1P_C8), e, splat2(DMFG_L1P_C7));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C6));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C5));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C4));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C3));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C2));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C1));
    pl = __ffma2_rn(pl, e, splat2(DMFG_L1P_C0));
    return __fmul2_rn(pl, e);
}
// packed (two elements per instruction) forms of the two functions above
__device__ __forceinline__ void policy_alpha_fast2(float theta, float2 x, float2& alpha, float2& alpha_deriv) {
    const float2 t = __fmul2_rn(splat2(theta), x);
    const float2 arg = __fmul2_rn(make_float2(fabsf(t.x), fabsf(t.y)), splat2(-DMFG_LOG2E));
    const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
    const float2 u = __fadd2_rn(e, splat2(1.0f));
    const float2 l = log1p_poly2(e);
    const float2 ru = make_float2(rcp_approx(u.x), rcp_approx(u.y));
    const float2 sg = __fmul2_rn(make_float2(t.x >= 0.0f ? 1.0f : e.x, t.y >= 0.0f ? 1.0f : e.y), ru);
    alpha = __fadd2_rn(make_float2(fmaxf(t.x, 0.0f), fmaxf(t.y, 0.0f)), l);
    alpha_deriv = __fmul2_rn(x, sg);
}
__device__ __forceinline__ float2 digamma_fast2(float2 x) {
    float2 q = __ffma2_rn(x, x, __fmul2_rn(splat2(3.0f), x));
    q = make_float2(fmaxf(q.x, 1e-30f), fmaxf(q.y, 1e-30f));
    const float2 num = __fmul2_rn(__ffma2_rn(splat2(2.0f), x, splat2(3.0f)), __ffma2_rn(splat2(2.0f), q, splat2(2.0f)));
    const float2 den = __fmul2_rn(q, __fadd2_rn(q, splat2(2.0f)));
    const float2 z = __fadd2_rn(x, splat2(4.0f));
    const float2 dz = __fmul2_rn(den, z);
    const float2 R = make_float2(rcp_approx(dz.x), rcp_approx(dz.y));
    const float2 s = __fmul2_rn(num, __fmul2_rn(z, R));
    const float2 rz = __fmul2_rn(den, R);
    const float2 r2 = __fmul2_rn(rz, rz);
    float2 p = __ffma2_rn(r2, splat2(-1.0f / 252.0f), splat2(1.0f / 120.0f));
    p = __ffma2_rn(r2, p, splat2(-1.0f / 12.0f));
    float2 res = __ffma2_rn(make_float2(lg2_approx(z.x), lg2_approx(z.y)), splat2(DMFG_LN2), neg2(s));
    res = __ffma2_rn(splat2(-0.5f), rz, res);
    return __ffma2_rn(r2, p, res);
}

// alpha, alpha' and psi(alpha) of a pair in one go: the sigmoid's 1/(1+e) and the digamma's 1/(den z) share
// ONE reciprocal (R0 = 1/((1+e) den z)), so a pair costs 12 MUFU operations instead of 14
__device__ __forceinline__ void alpha_psi_fast2(float theta, float2 x, float2& alpha, float2& alpha_deriv, float2& psi) {
    const float2 t = __fmul2_rn(splat2(theta), x);
    const float2 arg = __fmul2_rn(make_float2(fabsf(t.x), fabsf(t.y)), splat2(-DMFG_LOG2E));
    const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
    const float2 u = __fadd2_rn(e, splat2(1.0f));
    const float2 l = log1p_poly2(e);
    // max(t, 0) + l as 0.5 (t + |t|) + l: t + |t| and the halving are exact, so the bits equal the scalar form
    const float2 a = __ffma2_rn(__fadd2_rn(t, make_float2(fabsf(t.x), fabsf(t.y))), splat2(0.5f), l);
    alpha = a;
    // digamma(a): 4-step recurrence folded into one quotient + asymptotic series at z = a + 4
    // (q = a^2 + 3a, kept away from 0 by a 1e-30 addend that vanishes in the rounding for any a > 1e-22)
    const float2 q = __ffma2_rn(a, a, __ffma2_rn(splat2(3.0f), a, splat2(1e-30f)));
    const float2 num = __fmul2_rn(__ffma2_rn(splat2(2.0f), a, splat2(3.0f)), __ffma2_rn(splat2(2.0f), q, splat2(2.0f)));
    const float2 den = __fmul2_rn(q, __fadd2_rn(q, splat2(2.0f)));
    const float2 z = __fadd2_rn(a, splat2(4.0f));
    const float2 dz = __fmul2_rn(den, z);
    const float2 m = __fmul2_rn(u, dz);
    const float2 R0 = make_float2(rcp_approx(m.x), rcp_approx(m.y));
    const float2 ru = __fmul2_rn(dz, R0);                       // 1 / (1 + e)
    const float2 R = __fmul2_rn(u, R0);                         // 1 / (den z)
    const float2 sg = __fmul2_rn(make_float2(t.x >= 0.0f ? 1.0f : e.x, t.y >= 0.0f ? 1.0f : e.y), ru);
    alpha_deriv = __fmul2_rn(x, sg);
    const float2 srec = __fmul2_rn(num, __fmul2_rn(z, R));
    const float2 rz = __fmul2_rn(den, R);
    const float2 r2 = __fmul2_rn(rz, rz);
    float2 p = __ffma2_rn(r2, splat2(-1.0f / 252.0f), splat2(1.0f / 120.0f));
    p = __ffma2_rn(r2, p, splat2(-1.0f / 12.0f));
    float2 res = __ffma2_rn(make_float2(lg2_approx(z.x), lg2_approx(z.y)), splat2(DMFG_LN2), neg2(srec));
    res = __ffma2_rn(splat2(-0.5f), rz, res);
    psi = __ffma2_rn(r2, p, res);
}

// ------------------------------------------------------- sub-warp reductions
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, G);
    return v;
}

}  // namespace dmfg
