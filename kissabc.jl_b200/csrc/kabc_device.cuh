// kabc_device.cuh -- device building blocks of the KissABC hot path (sm_100a).
//   * Philox4x32-10 counter streams (DESIGN.md "Variate spec")
//   * exactly-rounded FP64 elementary functions: a fixed sequence of IEEE ops, so the F64 path is
//     bit-reproducible on any IEEE machine (this is what the parity tests rely on)
//   * priors: logpdf(Factored) ref src/priors.jl:30-36, rand(Factored) ref src/priors.jl:42-43
//   * fast FP32 Box-Muller on the MUFU pipe for the F32_ACC64 simulators
// Compile with -fmad=false: every fused multiply-add in this file is explicit.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/kissabc_cuda.h"

namespace kabc {

enum StreamTag : uint32_t { ST_PRIOR = 1, ST_PROPOSE = 2, ST_COST = 3, ST_ACCEPT = 4, ST_COST_INIT = 5 };

// Philox round keys, precomputed on the host: they are launch constants, so in SASS they become
// constant-bank operands of the LOP3 and cost no instruction.
struct RoundKeys {
    uint32_t k0[10], k1[10];
};

__host__ __device__ inline RoundKeys make_round_keys(uint64_t seed) {
    RoundKeys rk;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk.k0[r] = a;
        rk.k1[r] = b;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    return rk;
}

__device__ __forceinline__ void philox4x32_10(const RoundKeys &rk, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk.k0[r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// sequential word stream (seed ; block j, id, epoch, tag): used by the low-rate code (proposals, priors, accept)
struct Stream {
    const RoundKeys &rk;
    uint32_t j, id, epoch, tag;
    uint32_t b0, b1, b2, b3;
    int pos;
    __device__ __forceinline__ Stream(const RoundKeys &rk_, uint32_t tag_, uint32_t id_, uint32_t epoch_)
        : rk(rk_), j(0), id(id_), epoch(epoch_), tag(tag_), b0(0), b1(0), b2(0), b3(0), pos(4) {}
    __device__ __forceinline__ uint32_t next() {
        if (pos == 4) {
            philox4x32_10(rk, j, id, epoch, tag, b0, b1, b2, b3);
            j += 1;
            pos = 0;
        }
        uint32_t w = pos == 0 ? b0 : (pos == 1 ? b1 : (pos == 2 ? b2 : b3));
        pos += 1;
        return w;
    }
};

// ---------------------------------------------------------------- exact FP64 functions
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xfma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7FF0000000000000ll); }
__device__ __forceinline__ double dnan() { return __longlong_as_double(0x7FF8000000000000ll); }
__device__ __forceinline__ bool dfinite(double x) {
    return ((unsigned long long)__double_as_longlong(x) & 0x7FF0000000000000ull) != 0x7FF0000000000000ull;
}

// Polynomial coefficients live in constant memory so that DFMA reads them as constant-bank operands (as 64-bit
// immediates they cost two MOVs each).  The initialisers are folded by the host compiler with IEEE division, i.e.
// they are the same correctly rounded doubles the oracle's literals are.
static __constant__ double KC_LOG[11] = {1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0,
                                         1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};
static __constant__ double KC_EXP[13] = {1.0 / 87178291200.0, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0,
                                         1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0,
                                         1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
static __constant__ double KC_SIN[9] = {-1.0 / 121645100408832000.0, 1.0 / 355687428096000.0, -1.0 / 1307674368000.0,
                                        1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0,
                                        -1.0 / 6.0};
static __constant__ double KC_COS[9] = {1.0 / 6402373705728000.0, -1.0 / 20922789888000.0, 1.0 / 87178291200.0,
                                        -1.0 / 479001600.0, 1.0 / 3628800.0, -1.0 / 40320.0, 1.0 / 720.0, -1.0 / 24.0, 0.5};
static __constant__ double KC_MISC[6] = {6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.4426950408889634,
                                         6.283185307179586, 1.4142135623730951, 2.3283064365386962890625e-10};

static __device__ __noinline__ double xlog(double x) {
    if (x != x) return x;
    if (x < 0.0) return dnan();
    if (x == 0.0) return -dinf();
    if (x == dinf()) return x;
    int e = 0;
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    if ((b >> 52) == 0) {
        x = xmul(x, 18014398509481984.0);
        b = (unsigned long long)__double_as_longlong(x);
        e = -54;
    }
    e += (int)(b >> 52) - 1023;
    double m = __longlong_as_double((long long)((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull));
    if (m > KC_MISC[4]) { m = xmul(m, 0.5); e += 1; }
    double s = xdiv(xsub(m, 1.0), xadd(m, 1.0));
    double s2 = xmul(s, s);
    double p = KC_LOG[0];
#pragma unroll
    for (int q = 1; q < 11; ++q) p = xfma(p, s2, KC_LOG[q]);
    double t = xmul(xmul(s, s2), p);
    double r = xadd(xmul(2.0, s), xmul(2.0, t));
    double ef = (double)e;
    return xadd(xmul(ef, KC_MISC[0]), xadd(r, xmul(ef, KC_MISC[1])));
}

static __device__ __noinline__ double xexp(double x) {
    if (x != x) return x;
    if (x > 709.78) return dinf();
    if (x < -745.2) return 0.0;
    double k = floor(xadd(xmul(x, KC_MISC[2]), 0.5));
    double r = xfma(-k, KC_MISC[0], x);
    r = xfma(-k, KC_MISC[1], r);
    double p = KC_EXP[0];
#pragma unroll
    for (int q = 1; q < 13; ++q) p = xfma(p, r, KC_EXP[q]);
    p = xfma(p, r, 1.0);
    p = xfma(p, r, 1.0);
    int ki = (int)k;
    int k1 = ki / 2, k2 = ki - k1;
    double f1 = __longlong_as_double((long long)(k1 + 1023) << 52);
    double f2 = __longlong_as_double((long long)(k2 + 1023) << 52);
    return xmul(xmul(p, f1), f2);
}

__device__ __forceinline__ void xsincos2pi(double u, double &sn, double &cs) {
    double q = floor(xadd(xmul(4.0, u), 0.5));
    double t = xsub(u, xmul(0.25, q));
    double phi = xmul(t, KC_MISC[3]);
    double p2 = xmul(phi, phi);
    double ps = KC_SIN[0], pc = KC_COS[0];
#pragma unroll
    for (int j = 1; j < 9; ++j) {
        ps = xfma(ps, p2, KC_SIN[j]);
        pc = xfma(pc, p2, KC_COS[j]);
    }
    double s = xfma(xmul(phi, p2), ps, phi);
    double c = xfma(-p2, pc, 1.0);
    int qi = ((int)q) & 3;
    sn = qi == 0 ? s : (qi == 1 ? c : (qi == 2 ? -s : -c));
    cs = qi == 0 ? c : (qi == 1 ? -s : (qi == 2 ? -c : s));
}

__device__ __forceinline__ double u01(uint32_t w) { return xmul(xadd((double)w, 0.5), KC_MISC[5]); }
__device__ __forceinline__ uint32_t index_of(uint32_t w, uint32_t n) { return __umulhi(w, n); }

__device__ __forceinline__ void normal_pair64(uint32_t w0, uint32_t w1, double &z0, double &z1) {
    double r = xsqrt(xmul(-2.0, xlog(u01(w0))));
    double s, c;
    xsincos2pi(u01(w1), s, c);
    z0 = xmul(r, c);
    z1 = xmul(r, s);
}
__device__ __forceinline__ double next_uniform(Stream &st) { return u01(st.next()); }
__device__ __forceinline__ double next_normal(Stream &st) {
    uint32_t w0 = st.next(), w1 = st.next();
    double z0, z1;
    normal_pair64(w0, w1, z0, z1);
    return z0;
}
__device__ __forceinline__ double next_exp(Stream &st) { return -xlog(next_uniform(st)); }

// ---------------------------------------------------------------- fast FP32 normals (MUFU pipe)
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Same words, same Box-Muller map as normal_pair64, evaluated in FP32 with hardware approximations:
// u = (w+0.5)/2^32 rounded to float; r = sqrt(-2 ln2 * lg2(u1)); angle = 2 pi u2.
__device__ __forceinline__ void normal_pair32(uint32_t w0, uint32_t w1, float &z0, float &z1) {
    float u1 = __fmaf_rn(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float ang = __fmaf_rn(__uint2float_rn(w1), 1.4629180792671596e-09f, 7.3145903963357980e-10f); // 2pi(w+0.5)/2^32
    float r = mufu_sqrt(__fmul_rn(mufu_lg2(u1), -1.3862943611198906f));
    z0 = __fmul_rn(r, mufu_cos(ang));
    z1 = __fmul_rn(r, mufu_sin(ang));
}

// The normal model only needs sum(z) and sum(z^2) of its draws, and for a Box-Muller pair both follow from the pair's
// polar form without forming z0, z1:  z0 + z1 = r (cos t + sin t) = sqrt(2) r sin(t + pi/4),  z0^2 + z1^2 = r^2 = -2 ln u1.
// With s = sqrt(-lg2 u1): z0 + z1 = 2 sqrt(ln 2) * s sin(t + pi/4) and z0^2 + z1^2 = -2 ln 2 * lg2 u1, so a pair costs
// three MUFU (lg2, sqrt, sin) + one FFMA + one FADD instead of four MUFU + 2 FMUL + 2 FADD + 2 FFMA + 1 FMUL; the constant
// factors are applied once per particle in FP64.  Same Philox words, same u1 / u2 as normal_pair32.
__device__ __forceinline__ void normal_pair_sums32(uint32_t w0, uint32_t w1, float &acc_s_sin, float &acc_lg2) {
    const float u1 = __fmaf_rn(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    // angle 2 pi u2 + pi/4 from the top 23 bits of w1 placed in the mantissa of a float in [1,2) (two ALU operations;
    // an I2F would go to the XU pipe, which the three MUFU of the pair already load as much as Philox loads IMAD.WIDE):
    // f = 1 + floor(w1 / 2^9) 2^-23, ang = 2 pi (f - 1 + 2^-24) + pi/4, |u2 - (w1 + 0.5) 2^-32| < 2^-24
    const float f = __uint_as_float((w1 >> 9) | 0x3f800000u);
    const float ang = __fmaf_rn(f, 6.2831853071795865f, -5.4977867692f);
    const float l = mufu_lg2(u1);
    acc_s_sin = __fmaf_rn(mufu_sqrt(-l), mufu_sin(ang), acc_s_sin);
    acc_lg2 = __fadd_rn(acc_lg2, l);
}

// log Gamma(x), x > 0 (part of the variate spec): recurrence up to x >= 10, then the Stirling series through x^-13
static __constant__ double KC_LGAMMA[7] = {1.0 / 156.0, -691.0 / 360360.0, 1.0 / 1188.0, -1.0 / 1680.0, 1.0 / 1260.0,
                                           -1.0 / 360.0, 1.0 / 12.0};
static __device__ __noinline__ double xlgamma(double x) {
    if (x != x) return x;
    if (x < 0.0) return dnan();
    if (x == 0.0 || x == dinf()) return dinf();
    double prod = 1.0;
    while (x < 10.0) { prod = xmul(prod, x); x = xadd(x, 1.0); }
    double xi = xdiv(1.0, x);
    double x2 = xmul(xi, xi);
    double p = KC_LGAMMA[0];
#pragma unroll
    for (int q = 1; q < 7; ++q) p = xfma(p, x2, KC_LGAMMA[q]);
    double r = xadd(xadd(xsub(xmul(xsub(x, 0.5), xlog(x)), x), 0.91893853320467274178), xmul(p, xi));
    return xsub(r, xlog(prod));
}

// ---------------------------------------------------------------- priors
#define KABC_LOG2PI 1.8378770664093453
struct DPrior {
    int kind;
    double p0, p1, lo, hi;
    // host libm, same expressions as the oracle:
    double c0; // Uniform: log(b-a); Normal/Truncated: log(sigma); Beta: logbeta(a,b); NegativeBinomial: lgamma(r);
               // DiscreteUniform: log(b-a+1)
    double c1; // Truncated: log(Phi((hi-mu)/sigma) - Phi((lo-mu)/sigma)); NegativeBinomial: r*log(p)
    double c2; // NegativeBinomial: log1p(-p)
};
struct DPriors {
    int d;
    DPrior p[KABC_MAX_DIM];
};
__host__ __device__ __forceinline__ bool prior_is_discrete(const DPrior &p) {
    return p.kind == KABC_PRIOR_NEG_BINOMIAL || p.kind == KABC_PRIOR_DISCRETE_UNIFORM;
}
// ref src/types.jl:28-32 push_p: round(Int, .) (ties to even) for discrete laws, identity otherwise
__device__ __forceinline__ double push1(const DPrior &p, double x) { return prior_is_discrete(p) ? rint(x) : x; }
__host__ __device__ __forceinline__ uint32_t push_mask_of(const DPriors &P) {
    uint32_t m = 0;
    for (int k = 0; k < P.d; ++k) m |= prior_is_discrete(P.p[k]) ? (1u << k) : 0u;
    return m;
}

// the laws outside the headline workloads stay out of line so the hot kernels only pay a compare for them
// (scalars, not a DPrior reference: taking the address of a kernel parameter would spill the whole DPriors to local memory)
static __device__ __noinline__ double prior1_logpdf_ext(int kind, double p0, double p1, double c0, double c1, double c2,
                                                        double x) {
    if (kind == KABC_PRIOR_BETA) {
        if (!(x >= 0.0 && x <= 1.0)) return -dinf();
        double am = xsub(p0, 1.0), bm = xsub(p1, 1.0);
        double t1 = am == 0.0 ? 0.0 : xmul(am, xlog(x));
        double t2 = bm == 0.0 ? 0.0 : xmul(bm, xlog(xsub(1.0, x)));
        return xsub(xadd(t1, t2), c0);
    }
    if (kind == KABC_PRIOR_NEG_BINOMIAL) {
        if (!(x >= 0.0) || x != floor(x) || x == dinf()) return -dinf();
        return xadd(xsub(xsub(xlgamma(xadd(x, p0)), xlgamma(xadd(x, 1.0))), c0), xadd(c1, xmul(x, c2)));
    }
    if (kind == KABC_PRIOR_DISCRETE_UNIFORM) return (x >= p0 && x <= p1 && x == floor(x)) ? -c0 : -dinf();
    return dnan();
}

__device__ __forceinline__ double prior1_logpdf(const DPrior &p, double x) {
    if (p.kind >= KABC_PRIOR_BETA) return prior1_logpdf_ext(p.kind, p.p0, p.p1, p.c0, p.c1, p.c2, x);
    if (p.kind == KABC_PRIOR_UNIFORM) return (x >= p.p0 && x <= p.p1) ? -p.c0 : -dinf();
    if (p.kind == KABC_PRIOR_TRUNC_NORMAL && !(x >= p.lo && x <= p.hi)) return -dinf();
    double z = xdiv(xsub(x, p.p0), p.p1);
    double base = xsub(xmul(-xadd(xmul(z, z), KABC_LOG2PI), 0.5), p.c0); // x * 0.5 == x / 2 exactly
    return p.kind == KABC_PRIOR_NORMAL ? base : xsub(base, p.c1);
}
// ref src/priors.jl:30-36: left-to-right sum from component 1
template <typename F>
__device__ __forceinline__ double prior_logpdf(const DPriors &P, F get) {
    double s = prior1_logpdf(P.p[0], get(0));
    for (int k = 1; k < P.d; ++k) s = xadd(s, prior1_logpdf(P.p[k], get(k)));
    return s;
}
// logpdf(prior, push_p(prior, x)), ref src/smc.jl:125,172 and src/KissABC.jl:51
template <typename F>
__device__ __forceinline__ double prior_logpdf_pushed(const DPriors &P, F get) {
    return prior_logpdf(P, [&](int k) { return push1(P.p[k], get(k)); });
}

// Gamma(shape a, 1), Marsaglia & Tsang; Poisson: Knuth below 10, PTRS from 10 up (part of the variate spec)
#define KABC_GAMMA_MAX_TRIES 4096
static __device__ __noinline__ bool gamma_sample(Stream &st, double a, double &out) {
    double boost = 1.0;
    if (a < 1.0) {
        boost = xexp(xdiv(xlog(next_uniform(st)), a));
        a = xadd(a, 1.0);
    }
    double d = xsub(a, 1.0 / 3.0);
    double c = xdiv(1.0, xsqrt(xmul(9.0, d)));
    for (int t = 0; t < KABC_GAMMA_MAX_TRIES; ++t) {
        double z = next_normal(st);
        double u = next_uniform(st);
        double v = xadd(1.0, xmul(c, z));
        if (!(v > 0.0)) continue;
        v = xmul(xmul(v, v), v);
        if (xlog(u) < xadd(xsub(xadd(xmul(xmul(0.5, z), z), d), xmul(d, v)), xmul(d, xlog(v)))) {
            out = xmul(xmul(d, v), boost);
            return true;
        }
    }
    out = dnan();
    return false;
}
static __device__ __noinline__ bool poisson_sample(Stream &st, double lam, double &out) {
    if (!(lam >= 0.0) || lam > 1e9) { out = dnan(); return false; }
    if (lam == 0.0) { out = 0.0; return true; }
    if (lam < 10.0) {
        double L = xexp(-lam), p = 1.0;
        for (int k = 0; k < KABC_GAMMA_MAX_TRIES; ++k) {
            p = xmul(p, next_uniform(st));
            if (!(p > L)) { out = (double)k; return true; }
        }
        out = dnan();
        return false;
    }
    double slam = xsqrt(lam), loglam = xlog(lam);
    double b = xadd(0.931, xmul(2.53, slam));
    double a = xadd(-0.059, xmul(0.02483, b));
    double invalpha = xadd(1.1239, xdiv(1.1328, xsub(b, 3.4)));
    double vr = xsub(0.9277, xdiv(3.6224, xsub(b, 2.0)));
    for (int t = 0; t < KABC_GAMMA_MAX_TRIES; ++t) {
        double U = xsub(next_uniform(st), 0.5);
        double V = next_uniform(st);
        double us = xsub(0.5, fabs(U));
        double k = floor(xadd(xadd(xmul(xadd(xdiv(xmul(2.0, a), us), b), U), lam), 0.43));
        if (us >= 0.07 && V <= vr) { out = k; return true; }
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        double lhs = xsub(xadd(xlog(V), xlog(invalpha)), xlog(xadd(xdiv(a, xmul(us, us)), b)));
        if (lhs <= xsub(xsub(xmul(k, loglam), lam), xlgamma(xadd(k, 1.0)))) { out = k; return true; }
    }
    out = dnan();
    return false;
}
static __device__ __noinline__ bool prior1_sample_ext(int kind, double p0, double p1, Stream &st, double &out) {
    if (kind == KABC_PRIOR_BETA) {
        double ga, gb;
        bool ok = gamma_sample(st, p0, ga);
        ok = gamma_sample(st, p1, gb) && ok;
        out = xdiv(ga, xadd(ga, gb));
        return ok;
    }
    if (kind == KABC_PRIOR_NEG_BINOMIAL) {
        double g;
        if (!gamma_sample(st, p0, g)) { out = dnan(); return false; }
        return poisson_sample(st, xmul(g, xdiv(xsub(1.0, p1), p1)), out);
    }
    out = xadd(p0, (double)index_of(st.next(), (uint32_t)xadd(xsub(p1, p0), 1.0)));
    return true;
}

#define KABC_TRUNC_MAX_TRIES (1 << 20)
__device__ __forceinline__ bool prior1_sample(const DPrior &p, Stream &st, double &out) {
    if (p.kind >= KABC_PRIOR_BETA) return prior1_sample_ext(p.kind, p.p0, p.p1, st, out);
    if (p.kind == KABC_PRIOR_UNIFORM) {
        out = xadd(p.p0, xmul(xsub(p.p1, p.p0), next_uniform(st)));
        return true;
    }
    if (p.kind == KABC_PRIOR_NORMAL) {
        out = xadd(p.p0, xmul(p.p1, next_normal(st)));
        return true;
    }
    for (int t = 0; t < KABC_TRUNC_MAX_TRIES; ++t) {
        double x = xadd(p.p0, xmul(p.p1, next_normal(st)));
        if (x >= p.lo && x <= p.hi) { out = x; return true; }
    }
    out = dnan();
    return false;
}

// ---------------------------------------------------------------- small reductions
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

} // namespace kabc
