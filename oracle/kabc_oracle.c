/*
 * kabc_oracle.c -- CPU ORACLE (test infrastructure, not product).  See kabc_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fma()
 * so the arithmetic is a fixed sequence of correctly rounded IEEE-754 double
 * operations that the device F64 path reproduces bit for bit.
 *
 * Citations "ref:" are into /root/reference (KissABC.jl 3.0.1).
 */
#include "kabc_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[256];
const char *kor_last_error(void) { return g_err; }
static int fail(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 1;
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11): counter-based RNG             */
/* ------------------------------------------------------------------ */
void kor_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream = (seed ; block j, id, epoch, stream-tag); words are consumed in order */
enum { ST_PRIOR = 1, ST_PROPOSE = 2, ST_COST = 3, ST_ACCEPT = 4, ST_COST_INIT = 5, ST_SERIAL = 6 };
/* SERIAL mode (kor_smc_set_serial / kor_ais_set_serial): every variate of a run comes from ONE word stream
 * (seed ; tag SERIAL, id 0, epoch 0), consumed in the order the REFERENCE consumes its single `rng` -- init draws of all
 * particles, then all initial costs, then per sweep all proposals in particle order, then all costs in particle order
 * (ref src/smc.jl:119-123,160-191).  julia/PhiloxRNG.jl implements the same word stream as an AbstractRNG, so that the
 * unmodified KissABC.jl can be run on it and compared with this mode (julia/make_ref_fixtures.jl, tests/test_ref_fixtures.py). */
typedef struct {
    uint32_t key[2], ctr[4], buf[4];
    int pos;
} stream_t;

static void stream_init(stream_t *s, uint64_t seed, uint32_t tag, uint32_t id, uint32_t epoch) {
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = 0; s->ctr[1] = id; s->ctr[2] = epoch; s->ctr[3] = tag;
    s->pos = 4;
}
static uint32_t next_u32(stream_t *s) {
    if (s->pos == 4) {
        kor_philox4x32_10(s->ctr, s->key, s->buf);
        s->ctr[0] += 1;
        s->pos = 0;
    }
    return s->buf[s->pos++];
}
uint32_t kor_stream_word(uint64_t seed, uint32_t tag, uint32_t id, uint32_t epoch, uint32_t k) {
    stream_t s;
    stream_init(&s, seed, tag, id, epoch);
    s.ctr[0] = k / 4;
    uint32_t w = 0;
    for (uint32_t j = 0; j <= k % 4; ++j) w = next_u32(&s);
    return w;
}

/* ------------------------------------------------------------------ */
/* exactly reproducible elementary functions (DESIGN.md "Variate spec") */
/* only +,-,*,/,sqrt,floor and explicit fma: same bits on CPU and GPU   */
/* ------------------------------------------------------------------ */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

double kor_log(double x) {
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    int e = 0;
    uint64_t b = d2u(x);
    if ((b >> 52) == 0) { /* subnormal: scale by 2^54 */
        x = x * 18014398509481984.0;
        b = d2u(x);
        e = -54;
    }
    e += (int)(b >> 52) - 1023;
    double m = u2d((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    /* log(m) = 2 atanh(s) = 2s + 2 s^3 (1/3 + s2/5 + ... + s2^10/23) */
    double p = 1.0 / 23.0;
    p = fma(p, s2, 1.0 / 21.0);
    p = fma(p, s2, 1.0 / 19.0);
    p = fma(p, s2, 1.0 / 17.0);
    p = fma(p, s2, 1.0 / 15.0);
    p = fma(p, s2, 1.0 / 13.0);
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    double t = (s * s2) * p;
    double r = 2.0 * s + 2.0 * t;
    double ef = (double)e;
    /* ln2 split: hi has 21 trailing zero bits so ef*hi is exact */
    return ef * 6.93147180369123816490e-01 + (r + ef * 1.90821492927058770002e-10);
}

double kor_exp(double x) {
    if (x != x) return x;
    if (x > 709.78) return INFINITY;
    if (x < -745.2) return 0.0;
    double k = floor(x * 1.4426950408889634 + 0.5);
    double r = fma(-k, 6.93147180369123816490e-01, x);
    r = fma(-k, 1.90821492927058770002e-10, r);
    double p = 1.0 / 87178291200.0; /* 1/14! */
    p = fma(p, r, 1.0 / 6227020800.0);
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int ki = (int)k;
    /* 2^ki in two exact factors so results stay finite/normal whenever representable */
    int k1 = ki / 2, k2 = ki - k1;
    double f1 = u2d((uint64_t)(k1 + 1023) << 52);
    double f2 = u2d((uint64_t)(k2 + 1023) << 52);
    return (p * f1) * f2;
}

void kor_sincos2pi(double u, double *sn, double *cs) {
    double q = floor(4.0 * u + 0.5);
    double t = u - 0.25 * q; /* exact for the u grid we use */
    double phi = t * 6.283185307179586;
    double p2 = phi * phi;
    double ps = -1.0 / 121645100408832000.0; /* -1/19! */
    ps = fma(ps, p2, 1.0 / 355687428096000.0); /* 1/17! */
    ps = fma(ps, p2, -1.0 / 1307674368000.0);  /* -1/15! */
    ps = fma(ps, p2, 1.0 / 6227020800.0);      /* 1/13! */
    ps = fma(ps, p2, -1.0 / 39916800.0);       /* -1/11! */
    ps = fma(ps, p2, 1.0 / 362880.0);          /* 1/9! */
    ps = fma(ps, p2, -1.0 / 5040.0);           /* -1/7! */
    ps = fma(ps, p2, 1.0 / 120.0);             /* 1/5! */
    ps = fma(ps, p2, -1.0 / 6.0);              /* -1/3! */
    double s = fma(phi * p2, ps, phi);
    double pc = 1.0 / 6402373705728000.0; /* 1/18! */
    pc = fma(pc, p2, -1.0 / 20922789888000.0); /* -1/16! */
    pc = fma(pc, p2, 1.0 / 87178291200.0);     /* 1/14! */
    pc = fma(pc, p2, -1.0 / 479001600.0);      /* -1/12! */
    pc = fma(pc, p2, 1.0 / 3628800.0);         /* 1/10! */
    pc = fma(pc, p2, -1.0 / 40320.0);          /* -1/8! */
    pc = fma(pc, p2, 1.0 / 720.0);             /* 1/6! */
    pc = fma(pc, p2, -1.0 / 24.0);             /* -1/4! */
    pc = fma(pc, p2, 0.5);                     /* 1/2! (sign applied below) */
    double c = fma(-p2, pc, 1.0);
    int qi = ((int)q) & 3;
    switch (qi) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
    }
}

double kor_u01(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386962890625e-10; }
uint32_t kor_index(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * (uint64_t)n) >> 32); }
void kor_normal_pair(uint32_t w0, uint32_t w1, double *z0, double *z1) {
    double r = sqrt(-2.0 * kor_log(kor_u01(w0)));
    double s, c;
    kor_sincos2pi(kor_u01(w1), &s, &c);
    *z0 = r * c;
    *z1 = r * s;
}
static double next_uniform(stream_t *s) { return kor_u01(next_u32(s)); }
static double next_normal(stream_t *s) { /* consumes 2 words, keeps the cosine branch */
    uint32_t w0 = next_u32(s), w1 = next_u32(s);
    double z0, z1;
    kor_normal_pair(w0, w1, &z0, &z1);
    return z0;
}
static double next_exp(stream_t *s) { return -kor_log(next_uniform(s)); }

/* log Gamma(x), x > 0: argument shifted to >= 10 by the recurrence, then the Stirling series (Abramowitz & Stegun 6.1.41)
 * through x^-13.  Part of the variate spec: fixed operation order, same bits on CPU and GPU. */
double kor_lgamma(double x) {
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0 || x == INFINITY) return INFINITY;
    double prod = 1.0;
    while (x < 10.0) { prod = prod * x; x = x + 1.0; }
    double xi = 1.0 / x;
    double x2 = xi * xi;
    double p = 1.0 / 156.0;
    p = fma(p, x2, -691.0 / 360360.0);
    p = fma(p, x2, 1.0 / 1188.0);
    p = fma(p, x2, -1.0 / 1680.0);
    p = fma(p, x2, 1.0 / 1260.0);
    p = fma(p, x2, -1.0 / 360.0);
    p = fma(p, x2, 1.0 / 12.0);
    double r = (((x - 0.5) * kor_log(x) - x) + 0.91893853320467274178) + p * xi;
    return r - kor_log(prod);
}

/* Gamma(shape a, scale 1), Marsaglia & Tsang (2000); a < 1 through Gamma(a+1) * U^(1/a).  Every try consumes one
 * normal (2 words) and one uniform (1 word). */
#define GAMMA_MAX_TRIES 4096
static int gamma_sample(stream_t *s, double a, double *out) {
    double boost = 1.0;
    if (a < 1.0) {
        boost = kor_exp(kor_log(next_uniform(s)) / a);
        a = a + 1.0;
    }
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    for (int t = 0; t < GAMMA_MAX_TRIES; ++t) {
        double z = next_normal(s);
        double u = next_uniform(s);
        double v = 1.0 + c * z;
        if (!(v > 0.0)) continue;
        v = (v * v) * v;
        if (kor_log(u) < (((0.5 * z) * z + d) - d * v) + d * kor_log(v)) {
            *out = (d * v) * boost;
            return 0;
        }
    }
    *out = NAN;
    return 1;
}
/* Poisson(lam): product-of-uniforms (Knuth) below 10, PTRS transformed rejection (Hoermann 1993) from 10 up. */
#define POISSON_MAX_TRIES 4096
static int poisson_sample(stream_t *s, double lam, double *out) {
    if (!(lam >= 0.0) || lam > 1e9) { *out = NAN; return 1; }
    if (lam == 0.0) { *out = 0.0; return 0; }
    if (lam < 10.0) {
        double L = kor_exp(-lam), p = 1.0;
        for (int k = 0; k < POISSON_MAX_TRIES; ++k) {
            p = p * next_uniform(s);
            if (!(p > L)) { *out = (double)k; return 0; }
        }
        *out = NAN;
        return 1;
    }
    double slam = sqrt(lam), loglam = kor_log(lam);
    double b = 0.931 + 2.53 * slam;
    double a = -0.059 + 0.02483 * b;
    double invalpha = 1.1239 + 1.1328 / (b - 3.4);
    double vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int t = 0; t < POISSON_MAX_TRIES; ++t) {
        double U = next_uniform(s) - 0.5;
        double V = next_uniform(s);
        double us = 0.5 - fabs(U);
        double k = floor(((2.0 * a) / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) { *out = k; return 0; }
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (kor_log(V) + kor_log(invalpha) - kor_log(a / (us * us) + b) <= (k * loglam - lam) - kor_lgamma(k + 1.0)) {
            *out = k;
            return 0;
        }
    }
    *out = NAN;
    return 1;
}

/* ------------------------------------------------------------------ */
/* priors -- ref: src/priors.jl:30-43 + Distributions.jl closed forms   */
/* ------------------------------------------------------------------ */
#define LOG2PI 1.8378770664093453
static double std_normal_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

static double prior1_logpdf(const kor_prior_t *p, double x) {
    switch (p->kind) {
    case KOR_PRIOR_UNIFORM: /* Distributions: insupport ? -log(b-a) : -Inf */
        return (x >= p->p0 && x <= p->p1) ? -log(p->p1 - p->p0) : -INFINITY;
    case KOR_PRIOR_NORMAL: {
        double z = (x - p->p0) / p->p1;
        return -(z * z + LOG2PI) / 2.0 - log(p->p1);
    }
    case KOR_PRIOR_TRUNC_NORMAL: {
        if (!(x >= p->lo && x <= p->hi)) return -INFINITY;
        double z = (x - p->p0) / p->p1;
        double logtp = log(std_normal_cdf((p->hi - p->p0) / p->p1) - std_normal_cdf((p->lo - p->p0) / p->p1));
        return (-(z * z + LOG2PI) / 2.0 - log(p->p1)) - logtp;
    }
    case KOR_PRIOR_BETA: { /* Distributions: xlogy(a-1,x) + xlog1py(b-1,-x) - logbeta(a,b) on [0,1] */
        if (!(x >= 0.0 && x <= 1.0)) return -INFINITY;
        double am = p->p0 - 1.0, bm = p->p1 - 1.0;
        double t1 = am == 0.0 ? 0.0 : am * kor_log(x);
        double t2 = bm == 0.0 ? 0.0 : bm * kor_log(1.0 - x);
        return (t1 + t2) - ((lgamma(p->p0) + lgamma(p->p1)) - lgamma(p->p0 + p->p1));
    }
    case KOR_PRIOR_NEG_BINOMIAL: { /* pmf(k) = Gamma(k+r)/(k! Gamma(r)) p^r (1-p)^k, k = 0,1,2,... */
        if (!(x >= 0.0) || x != floor(x) || x == INFINITY) return -INFINITY;
        double r = p->p0;
        return ((kor_lgamma(x + r) - kor_lgamma(x + 1.0)) - lgamma(r)) + (r * log(p->p1) + x * log1p(-p->p1));
    }
    case KOR_PRIOR_DISCRETE_UNIFORM:
        return (x >= p->p0 && x <= p->p1 && x == floor(x)) ? -log((p->p1 - p->p0) + 1.0) : -INFINITY;
    }
    return NAN;
}
/* ref: src/types.jl:28-32 -- push_p: components of a discrete law are rounded (round(Int, .), ties to even) before the prior
 * density and the cost see them; the stored particle keeps its real value */
static int prior_is_discrete(const kor_prior_t *p) {
    return p->kind == KOR_PRIOR_NEG_BINOMIAL || p->kind == KOR_PRIOR_DISCRETE_UNIFORM;
}
void kor_push_p(const kor_prior_t *prior, int d, const double *x, double *out) {
    for (int k = 0; k < d; ++k) out[k] = prior_is_discrete(&prior[k]) ? nearbyint(x[k]) : x[k];
}
/* ref: src/priors.jl:30-36 -- left-to-right sum starting from component 1 */
double kor_prior_logpdf(const kor_prior_t *prior, int d, const double *x) {
    double s = prior1_logpdf(&prior[0], x[0]);
    for (int k = 1; k < d; ++k) s += prior1_logpdf(&prior[k], x[k]);
    return s;
}
#define TRUNC_MAX_TRIES (1 << 20)
static int prior1_sample(const kor_prior_t *p, stream_t *s, double *out) {
    switch (p->kind) {
    case KOR_PRIOR_UNIFORM: *out = p->p0 + (p->p1 - p->p0) * next_uniform(s); return 0;
    case KOR_PRIOR_NORMAL: *out = p->p0 + p->p1 * next_normal(s); return 0;
    case KOR_PRIOR_TRUNC_NORMAL:
        for (int t = 0; t < TRUNC_MAX_TRIES; ++t) {
            double x = p->p0 + p->p1 * next_normal(s);
            if (x >= p->lo && x <= p->hi) { *out = x; return 0; }
        }
        *out = NAN;
        return 1;
    case KOR_PRIOR_BETA: { /* X = Ga/(Ga+Gb) */
        double ga, gb;
        int rc = gamma_sample(s, p->p0, &ga);
        rc |= gamma_sample(s, p->p1, &gb);
        *out = ga / (ga + gb);
        return rc;
    }
    case KOR_PRIOR_NEG_BINOMIAL: { /* Gamma(r, (1-p)/p) mixture of Poissons */
        double g;
        if (gamma_sample(s, p->p0, &g)) { *out = NAN; return 1; }
        return poisson_sample(s, g * ((1.0 - p->p1) / p->p1), out);
    }
    case KOR_PRIOR_DISCRETE_UNIFORM:
        *out = p->p0 + (double)kor_index(next_u32(s), (uint32_t)((p->p1 - p->p0) + 1.0));
        return 0;
    }
    return 1;
}
/* ref: src/priors.jl:42-43 -- components drawn in order from one stream */
int kor_prior_sample(uint64_t seed, const kor_prior_t *prior, int d, uint32_t id, uint32_t epoch, double *x) {
    stream_t s;
    stream_init(&s, seed, ST_PRIOR, id, epoch);
    int rc = 0;
    for (int k = 0; k < d; ++k) rc |= prior1_sample(&prior[k], &s, &x[k]);
    return rc;
}

/* ------------------------------------------------------------------ */
/* simulators + distances                      */
/* ------------------------------------------------------------------ */
static __thread int64_t g_last_events;
int64_t kor_last_events(void) { return g_last_events; }

/* ref: README.md:35-52 / test/runtests.jl:281-286.
 * x = randn(n).*sigma .+ mu ; hypot(mean(x)-t0, (std(x)-t1)*w), std with n-1.
 * Two sequential passes over the same counter-generated draws. */
static double cost_normal(const kor_model_t *m, stream_t *s0, const double *th) {
    const int n = m->n_draws;
    const double mu = th[0], sigma = th[1];
    double z[4];
    double stackbuf[1024];
    double *x = n <= 1024 ? stackbuf : (double *)malloc(sizeof(double) * (size_t)n);
    stream_t *s = s0; /* the draws are consumed once and kept (two passes over x[]) */
    double sum = 0.0;
    for (int j = 0; j < n; j += 4) {
        uint32_t w0 = next_u32(s), w1 = next_u32(s), w2 = next_u32(s), w3 = next_u32(s);
        kor_normal_pair(w0, w1, &z[0], &z[1]);
        kor_normal_pair(w2, w3, &z[2], &z[3]);
        for (int q = 0; q < 4 && j + q < n; ++q) {
            x[j + q] = z[q] * sigma + mu;
            sum += x[j + q];
        }
    }
    double mean = sum / (double)n;
    double ss = 0.0;
    for (int j = 0; j < n; ++j) {
        double dx = x[j] - mean;
        ss += dx * dx;
    }
    if (x != stackbuf) free(x);
    double sd = sqrt(ss / (double)(n - 1));
    double d1 = mean - m->target[0];
    double d2 = (sd - m->target[1]) * m->param[0];
    return sqrt(d1 * d1 + d2 * d2);
}

/* MA(2): y_t = e_{t+2} + th1 e_{t+1} + th2 e_t, t = 0..n-1 (n+2 normals);
 * tau_j = (1/n) sum_{t>=j} y_t y_{t-j}; distance = || tau - target ||_2;
 * +Inf outside the invertibility triangle (no variates consumed). */
static double cost_ma2(const kor_model_t *m, stream_t *s, const double *th) {
    const int n = m->n_draws;
    const double t1 = th[0], t2 = th[1];
    if (!(t1 > -2.0 && t1 < 2.0 && t1 + t2 > -1.0 && t1 - t2 < 1.0)) return INFINITY;
    double e0 = 0, e1 = 0, y1 = 0, y2 = 0, a1 = 0, a2 = 0;
    double z[4];
    int t = -2; /* index of the y produced by the current normal */
    for (int j = 0; j < n + 2; j += 4) {
        uint32_t w0 = next_u32(s), w1 = next_u32(s), w2 = next_u32(s), w3 = next_u32(s);
        kor_normal_pair(w0, w1, &z[0], &z[1]);
        kor_normal_pair(w2, w3, &z[2], &z[3]);
        for (int q = 0; q < 4 && j + q < n + 2; ++q, ++t) {
            double e2 = z[q];
            if (t >= 0) {
                double y = (e2 + t1 * e1) + t2 * e0;
                if (t >= 1) a1 += y * y1;
                if (t >= 2) a2 += y * y2;
                y2 = y1; y1 = y;
            }
            e0 = e1; e1 = e2;
        }
    }
    double d1 = a1 / (double)n - m->target[0];
    double d2 = a2 / (double)n - m->target[1];
    return sqrt(d1 * d1 + d2 * d2);
}

/* g-and-k: x = A + B (1 + c (1-e^{-gz})/(1+e^{-gz})) (1+z^2)^k z, c = param[0] (0.8);
 * summaries = order statistics at 1-based ranks round(i n/8), i=1..7; Euclidean distance. */
static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
static double gk_transform(double A, double B, double g, double k, double c, double z) {
    double eg = kor_exp(-g * z);
    double skew = 1.0 + c * ((1.0 - eg) / (1.0 + eg));
    double kurt = kor_exp(k * kor_log(1.0 + z * z));
    return A + ((B * skew) * kurt) * z;
}
static double cost_gk(const kor_model_t *m, stream_t *s, const double *th) {
    const int n = m->n_draws;
    double *x = (double *)malloc(sizeof(double) * (size_t)(n + 4));
    double z[4];
    for (int j = 0; j < n; j += 4) {
        uint32_t w0 = next_u32(s), w1 = next_u32(s), w2 = next_u32(s), w3 = next_u32(s);
        kor_normal_pair(w0, w1, &z[0], &z[1]);
        kor_normal_pair(w2, w3, &z[2], &z[3]);
        for (int q = 0; q < 4 && j + q < n; ++q)
            x[j + q] = gk_transform(th[0], th[1], th[2], th[3], m->param[0], z[q]);
    }
    qsort(x, (size_t)n, sizeof(double), cmp_double);
    double acc = 0.0;
    for (int i = 1; i <= 7; ++i) {
        int64_t rank = ((int64_t)i * n + 4) / 8; /* round(i n/8), half up */
        if (rank < 1) rank = 1;
        double dq = x[rank - 1] - m->target[i - 1];
        acc += dq * dq;
    }
    free(x);
    return sqrt(acc);
}

/* Lotka-Volterra, Gillespie direct method.  theta = log rates; param = {X0, Y0, T, n_grid, max_events};
 * target = [X(t_1..t_G), Y(t_1..t_G)], t_g = g T/G.  RMS distance; +Inf if the event cap is hit. */
static double cost_lv(const kor_model_t *m, stream_t *s, const double *th) {
    const double c1 = kor_exp(th[0]), c2 = kor_exp(th[1]), c3 = kor_exp(th[2]);
    double X = m->param[0], Y = m->param[1];
    const double T = m->param[2];
    const int G = (int)m->param[3];
    const int64_t max_events = (int64_t)m->param[4];
    const double dt = T / (double)G;
    double t = 0.0, acc = 0.0;
    int g = 0;
    int64_t ev = 0;
    while (g < G) {
        double a1 = c1 * X, a2 = (c2 * X) * Y, a3 = c3 * Y;
        double a0 = (a1 + a2) + a3;
        double tn;
        uint32_t w0 = 0, w1 = 0;
        if (a0 > 0.0) {
            if (ev >= max_events) { g_last_events = ev; return INFINITY; }
            w0 = next_u32(s); w1 = next_u32(s);
            tn = t + (-kor_log(kor_u01(w0))) / a0;
        } else {
            tn = INFINITY;
        }
        while (g < G && (double)(g + 1) * dt <= tn) { /* record the pre-event state */
            double dx = X - m->target[g], dy = Y - m->target[G + g];
            acc += dx * dx;
            acc += dy * dy;
            ++g;
        }
        if (g >= G) break;
        double r = kor_u01(w1) * a0;
        if (r < a1) X += 1.0;
        else if (r < a1 + a2) { X -= 1.0; Y += 1.0; }
        else Y -= 1.0;
        t = tn;
        ++ev;
    }
    g_last_events = ev;
    return sqrt(acc / (double)(2 * G));
}

/* same simulator, returning the grid observations instead of the distance (used to make synthetic targets) */
int kor_lv_trajectory(const kor_model_t *m, uint64_t seed, const double *th, uint32_t id, uint32_t epoch, double *out) {
    const int G = (int)m->param[3];
    stream_t s;
    stream_init(&s, seed, ST_COST, id, epoch);
    const double c1 = kor_exp(th[0]), c2 = kor_exp(th[1]), c3 = kor_exp(th[2]);
    double X = m->param[0], Y = m->param[1];
    const double T = m->param[2];
    const int64_t max_events = (int64_t)m->param[4];
    const double dt = T / (double)G;
    double t = 0.0;
    int g = 0;
    int64_t ev = 0;
    while (g < G) {
        double a1 = c1 * X, a2 = (c2 * X) * Y, a3 = c3 * Y;
        double a0 = (a1 + a2) + a3, tn;
        uint32_t w0 = 0, w1 = 0;
        if (a0 > 0.0) {
            if (ev >= max_events) return 1;
            w0 = next_u32(&s); w1 = next_u32(&s);
            tn = t + (-kor_log(kor_u01(w0))) / a0;
        } else tn = INFINITY;
        while (g < G && (double)(g + 1) * dt <= tn) { out[g] = X; out[G + g] = Y; ++g; }
        if (g >= G) break;
        double r = kor_u01(w1) * a0;
        if (r < a1) X += 1.0;
        else if (r < a1 + a2) { X -= 1.0; Y += 1.0; }
        else Y -= 1.0;
        t = tn;
        ++ev;
    }
    return 0;
}

/* deterministic costs used by the reference's own tests:
 * param[0]=0: |th0^2 + 1 - 1.5|  (test/runtests.jl:77-86, sim(mu)=mu*mu+1)
 * param[0]=1: |th0 - 1.5|        (test/runtests.jl:177-182) */
static double cost_det(const kor_model_t *m, stream_t *s, const double *th) {
    if (m->param[0] == 0.0) return fabs((th[0] * th[0] + 1.0) - m->target[0]);
    if (m->param[0] == 2.0) /* test/runtests.jl:105-112: sim((n,du)) = (n*n+du)*(n+randn()*0.01); |sim - 5.5| */
        return fabs((th[0] * th[0] + th[1]) * (th[0] + next_normal(s) * m->param[1]) - m->target[0]);
    return fabs(th[0] - m->target[0]);
}

/* "Tiny Data, ABC and the Socks of Karl Broman", ref test/runtests.jl:34-44: th = (n_socks, prop_pairs);
 * n_pairs = round(prop * floor(n/2)), n_odd = n - 2 n_pairs; pick min(n, n_picked) socks without replacement from the sorted
 * list [1,1,2,2,...,n_pairs,n_pairs, n_pairs+1, ..., n_pairs+n_odd]; pairs = picked - unique, odds = unique - pairs;
 * cost = |pairs - t0| + |odds - t1|.  Spec of the draw: forward Fisher-Yates, pick j swaps position j with
 * j + index(word_j, n - j); only the first m positions and the <= m displaced tail entries are materialised. */
#define SOCKS_MAX_PICKED 32
static double cost_socks(const kor_model_t *m, stream_t *s, const double *th) {
    double nf = th[0], prop = th[1];
    int n_picked = (int)m->param[0];
    if (!(nf >= 0.0) || !(nf <= 1e9) || nf != floor(nf) || !(prop >= 0.0 && prop <= 1.0)) return INFINITY;
    int64_t n = (int64_t)nf;
    int64_t n_pairs = (int64_t)nearbyint(prop * floor(nf / 2.0));
    int mp = n < n_picked ? (int)n : n_picked;
    int64_t head[SOCKS_MAX_PICKED], tpos[SOCKS_MAX_PICKED], tval[SOCKS_MAX_PICKED];
    int nt = 0;
    for (int j = 0; j < mp; ++j) head[j] = j;
    for (int j = 0; j < mp; ++j) {
        int64_t r = j + (int64_t)kor_index(next_u32(s), (uint32_t)(n - j));
        if (r < mp) { int64_t t = head[j]; head[j] = head[r]; head[r] = t; }
        else {
            int q = 0;
            while (q < nt && tpos[q] != r) ++q;
            if (q == nt) { tpos[nt] = r; tval[nt] = r; ++nt; }
            int64_t t = head[j]; head[j] = tval[q]; tval[q] = t;
        }
    }
    int lu = 0;
    int64_t lab[SOCKS_MAX_PICKED];
    for (int j = 0; j < mp; ++j) {
        int64_t sidx = head[j];
        int64_t l = sidx < 2 * n_pairs ? sidx / 2 : sidx - n_pairs;
        int seen = 0;
        for (int q = 0; q < lu; ++q) seen |= (lab[q] == l);
        if (!seen) lab[lu++] = l;
    }
    double pairs = (double)(mp - lu), odds = (double)(lu - (mp - lu));
    return fabs(pairs - m->target[0]) + fabs(odds - m->target[1]);
}

static double cost_on_stream(const kor_model_t *m, stream_t *s, const double *th) {
    g_last_events = 0;
    switch (m->kind) {
    case KOR_MODEL_NORMAL_MEANSTD: return cost_normal(m, s, th);
    case KOR_MODEL_MA2_AUTOCOV: return cost_ma2(m, s, th);
    case KOR_MODEL_GK_OCTILE: return cost_gk(m, s, th);
    case KOR_MODEL_LV_SSA: return cost_lv(m, s, th);
    case KOR_MODEL_DETERMINISTIC: return cost_det(m, s, th);
    case KOR_MODEL_SOCKS: return cost_socks(m, s, th);
    }
    return NAN;
}
static double cost_dispatch(const kor_model_t *m, uint64_t seed, uint32_t tag, int d, const double *th,
                            uint32_t id, uint32_t epoch) {
    (void)d;
    stream_t s;
    stream_init(&s, seed, tag, id, epoch);
    return cost_on_stream(m, &s, th);
}
double kor_cost(const kor_model_t *model, uint64_t seed, int d, const double *theta, uint32_t id, uint32_t epoch) {
    return cost_dispatch(model, seed, ST_COST, d, theta, id, epoch);
}
void kor_eval_cost(const kor_model_t *model, uint64_t seed, int d, const double *theta_soa, int64_t n,
                   uint32_t first_id, uint32_t epoch, double *out, int nthreads) {
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t i = 0; i < n; ++i) {
        double th[16];
        for (int k = 0; k < d; ++k) th[k] = theta_soa[(int64_t)k * n + i];
        out[i] = cost_dispatch(model, seed, ST_COST, d, th, first_id + (uint32_t)i, epoch);
    }
}

/* ------------------------------------------------------------------ */
/* Statistics.quantile, type 7 (alpha=beta=1) -- ref: src/smc.jl:134    */
/* [dep, restated from the published Statistics.jl algorithm]           */
/* ------------------------------------------------------------------ */
static double quantile7_sorted(const double *v, int64_t n, double p) {
    double aleph = (double)n * p + (1.0 - p);
    int64_t j = (int64_t)aleph; /* trunc */
    if (j < 1) j = 1;
    if (j > n - 1) j = n - 1;
    double gam = aleph - (double)j;
    if (gam < 0.0) gam = 0.0;
    if (gam > 1.0) gam = 1.0;
    double a, b;
    if (n == 1) { a = v[0]; b = v[0]; }
    else { a = v[j - 1]; b = v[j]; }
    if (isfinite(a) && isfinite(b)) return a + gam * (b - a);
    return (1.0 - gam) * a + gam * b;
}
double kor_quantile7(const double *v, int64_t n, double p) {
    double *c = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(c, v, sizeof(double) * (size_t)n);
    qsort(c, (size_t)n, sizeof(double), cmp_double);
    double q = quantile7_sorted(c, n, p);
    free(c);
    return q;
}

/* ------------------------------------------------------------------ */
/* smc -- ref: src/smc.jl:92-206                                        */
/* ------------------------------------------------------------------ */
struct kor_smc {
    uint64_t seed;
    kor_prior_t prior[16];
    int d;
    kor_model_t model;
    kor_smc_config_t cfg;
    int nthreads;
    int64_t N;
    double *th, *X, *lpi;      /* th SoA: th[k*N+i] */
    uint8_t *alive;
    double *th2, *X2, *lpi2;   /* gather scratch */
    /* proposal / trace buffers */
    int64_t *ta, *tb;
    double *tz, *tlprob, *tlpip, *txp, *thp;
    uint8_t *tdec;
    double eps;
    int flag;
    int64_t iteration, n_alive, accepted, cost_evals, events, sweeps;
    double eps_prev;
    int resampled;
    uint32_t next_epoch;
    const double *override_xp;
    kor_smc_log_t *log;
    int64_t nlog, caplog;
    int serial;   /* 1: one word stream in the reference's consumption order (see ST_SERIAL) */
    stream_t ser;
};

int kor_smc_create(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model,
                   const kor_smc_config_t *cfg, int nthreads, kor_smc_t **out) {
    /* ref: src/smc.jl:107-118 argument checks, same order and messages */
    if (!(cfg->min_r_ess > 0)) return fail("min_r_ess must be > 0.");
    if (!(cfg->mcmc_retrys >= 0)) return fail("mcmc_retrys must be >= 0.");
    if (!(cfg->alpha > 0)) return fail("alpha must be > 0.");
    if (!(cfg->r_epstol >= 0)) return fail("r_epstol must be >= 0");
    if (!(cfg->mcmc_tol >= 0)) return fail("mcmc_tol must be >= 0");
    if (!(cfg->max_stretch > 1)) return fail("max_stretch must be > 1");
    if (d < 1 || d > 16) return fail("d out of range");
    double mn = cfg->alpha < cfg->min_r_ess ? cfg->alpha : cfg->min_r_ess;
    int64_t min_np = (int64_t)ceil(3.0 * (double)d / mn);
    if (cfg->nparticles < min_np) {
        snprintf(g_err, sizeof g_err, "nparticles must be >= %lld.", (long long)min_np);
        return 1;
    }
    kor_smc_t *s = (kor_smc_t *)calloc(1, sizeof *s);
    s->seed = seed;
    memcpy(s->prior, prior, sizeof(kor_prior_t) * (size_t)d);
    s->d = d;
    s->model = *model;
    s->cfg = *cfg;
    s->nthreads = nthreads > 0 ? nthreads : 1;
    int64_t N = s->N = cfg->nparticles;
    size_t nd = (size_t)N * (size_t)d;
    s->th = (double *)calloc(nd, 8); s->th2 = (double *)calloc(nd, 8); s->thp = (double *)calloc(nd, 8);
    s->X = (double *)calloc((size_t)N, 8); s->X2 = (double *)calloc((size_t)N, 8);
    s->lpi = (double *)calloc((size_t)N, 8); s->lpi2 = (double *)calloc((size_t)N, 8);
    s->alive = (uint8_t *)calloc((size_t)N, 1);
    s->ta = (int64_t *)calloc((size_t)N, 8); s->tb = (int64_t *)calloc((size_t)N, 8);
    s->tz = (double *)calloc((size_t)N, 8); s->tlprob = (double *)calloc((size_t)N, 8);
    s->tlpip = (double *)calloc((size_t)N, 8); s->txp = (double *)calloc((size_t)N, 8);
    s->tdec = (uint8_t *)calloc((size_t)N, 1);
    s->eps = INFINITY;
    *out = s;
    return 0;
}
void kor_smc_destroy(kor_smc_t *s) {
    if (!s) return;
    free(s->th); free(s->th2); free(s->thp); free(s->X); free(s->X2); free(s->lpi); free(s->lpi2);
    free(s->alive); free(s->ta); free(s->tb); free(s->tz); free(s->tlprob); free(s->tlpip); free(s->txp);
    free(s->tdec); free(s->log); free(s);
}
void kor_smc_set_cost_override(kor_smc_t *s, const double *xp) { s->override_xp = xp; }
void kor_smc_set_serial(kor_smc_t *s) {
    s->serial = 1;
    s->nthreads = 1;
    stream_init(&s->ser, s->seed, ST_SERIAL, 0, 0);
}

/* ref: src/smc.jl:119-129 */
int kor_smc_init(kor_smc_t *s) {
    const int64_t N = s->N;
    const int d = s->d;
    int bad = 0;
    int64_t events = 0;
    if (s->serial) { /* ref :119-125: all prior draws, then all costs, on the one stream */
        for (int64_t i = 0; i < N; ++i)
            for (int k = 0; k < d; ++k) bad |= prior1_sample(&s->prior[k], &s->ser, &s->th[(int64_t)k * N + i]);
        for (int64_t i = 0; i < N; ++i) {
            double th[16];
            for (int k = 0; k < d; ++k) th[k] = s->th[(int64_t)k * N + i];
            kor_push_p(s->prior, d, th, th);
            s->X[i] = cost_on_stream(&s->model, &s->ser, th);
            events += g_last_events;
            s->lpi[i] = kor_prior_logpdf(s->prior, d, th);
            s->alive[i] = 1;
        }
    }
#pragma omp parallel for schedule(dynamic, 64) num_threads(s->nthreads) reduction(| : bad) reduction(+ : events)
    for (int64_t i = s->serial ? N : 0; i < N; ++i) {
        double th[16];
        bad |= kor_prior_sample(s->seed, s->prior, d, (uint32_t)i, 0, th);
        for (int k = 0; k < d; ++k) s->th[(int64_t)k * N + i] = th[k];
        kor_push_p(s->prior, d, th, th); /* ref :122-125 cost(push_p(prior, .)), logpdf(prior, push_p(prior, .)) */
        s->X[i] = cost_dispatch(&s->model, s->seed, ST_COST_INIT, d, th, (uint32_t)i, 0);
        events += g_last_events;
        s->lpi[i] = kor_prior_logpdf(s->prior, d, th);
        s->alive[i] = 1;
    }
    if (bad) return fail("prior sampling failed (truncation too extreme)");
    s->eps = INFINITY;
    s->flag = 0;
    s->iteration = 0;
    s->n_alive = N;
    s->accepted = 0;
    s->cost_evals = N;
    s->events = events;
    s->next_epoch = 0;
    s->nlog = 0;
    return 0;
}

/* ref: src/smc.jl:160-191 -- one synchronous MCMC sweep with epoch e */
static void smc_sweep(kor_smc_t *s, uint32_t e, int64_t lo, int64_t hi, int64_t *out_acc, int64_t *out_evals, int64_t *out_events) {
    const int64_t N = s->N;
    const int d = s->d;
    const double eps = s->eps;
    const int flag = s->flag;
    const double sqNp = sqrt((double)d);
    /* phase A (ref :160-167): proposals from the pre-sweep ensemble */
#pragma omp parallel for schedule(static) num_threads(s->nthreads)
    for (int64_t i = lo; i < hi; ++i) {
        s->tdec[i] = 0;
        s->ta[i] = s->tb[i] = -1;
        s->tz[i] = s->tlprob[i] = s->tlpip[i] = s->txp[i] = NAN;
        for (int k = 0; k < d; ++k) s->thp[(int64_t)k * N + i] = NAN;
        if (!s->alive[i]) continue;
        stream_t st0, *stp = &st0;
        if (s->serial) stp = &s->ser;
        else stream_init(&st0, s->seed, ST_PROPOSE, (uint32_t)i, e);
#define st (*stp)
        int64_t a = i, b = i;
        while (a == i) a = kor_index(next_u32(&st), (uint32_t)N);
        while (b == i || b == a) b = kor_index(next_u32(&st), (uint32_t)N);
        double z = next_normal(&st);
        double sc = (s->cfg.max_stretch * z) / sqNp;
        for (int k = 0; k < d; ++k) {
            const double *t = s->th + (int64_t)k * N;
            s->thp[(int64_t)k * N + i] = t[i] + (t[b] - t[a]) * sc;
        }
        s->tlprob[i] = kor_log(next_uniform(&st));
#undef st
        s->ta[i] = a; s->tb[i] = b; s->tz[i] = z;
    }
    /* phase B (ref :168-191) */
    int64_t acc = 0, evals = 0, events = 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(s->nthreads) reduction(+ : acc, evals, events)
    for (int64_t i = lo; i < hi; ++i) {
        if (!s->alive[i]) continue;
        double thp[16];
        for (int k = 0; k < d; ++k) thp[k] = s->thp[(int64_t)k * N + i];
        kor_push_p(s->prior, d, thp, thp); /* ref :172,176 */
        double lpip = kor_prior_logpdf(s->prior, d, thp);
        s->tlpip[i] = lpip;
        if (lpip < 0 && !isfinite(lpip)) { s->tdec[i] = 1; continue; }
        double dl = (lpip - s->lpi[i]) + 0.0;
        double lM = dl != dl ? dl : fmin(dl, 0.0); /* Julia's min propagates NaN (C's fmin does not): `lprob < NaN` skips, ref :174-175 */
        if (!(s->tlprob[i] < lM)) { s->tdec[i] = 2; continue; }
        double Xp;
        if (s->override_xp) Xp = s->override_xp[i];
        else {
            Xp = s->serial ? cost_on_stream(&s->model, &s->ser, thp) : cost_dispatch(&s->model, s->seed, ST_COST, d, thp, (uint32_t)i, e);
            events += g_last_events;
        }
        evals += 1;
        s->txp[i] = Xp;
        if (flag ? (Xp > eps) : (Xp >= eps)) { s->tdec[i] = 3; continue; }
        for (int k = 0; k < d; ++k) s->th[(int64_t)k * N + i] = s->thp[(int64_t)k * N + i]; /* ref :182 stores the raw proposal */
        s->X[i] = Xp;
        s->lpi[i] = lpip;
        s->tdec[i] = 4;
        acc += 1;
    }
    *out_acc = acc;
    *out_evals = evals;
    *out_events = events;
}

/* the `while true` body in three parts so that a sharded (multi-rank) schedule can be emulated:
 * cut (ref :132-153) ; [sweep over a shard ; commit of the summed counters]* (ref :156-193) ; finish (ref :194-198) */
int kor_smc_cut(kor_smc_t *s) {
    const int64_t N = s->N;
    const int d = s->d;
    const kor_smc_config_t *c = &s->cfg;
    s->iteration += 1;
    s->eps_prev = s->eps;
    /* ref :134 quantile(Xs[alive], alpha) and :136 minimum(Xs[alive]) */
    int64_t na = 0;
    double *xa = s->X2;
    double mn = INFINITY;
    for (int64_t i = 0; i < N; ++i)
        if (s->alive[i]) { xa[na++] = s->X[i]; if (s->X[i] < mn) mn = s->X[i]; }
    if (na == 0) return fail("no alive particles");
    qsort(xa, (size_t)na, sizeof(double), cmp_double);
    double eps = quantile7_sorted(xa, na, c->alpha);
    int flag = 0;
    int64_t ess = 0;
    if (eps > mn) {
        for (int64_t i = 0; i < N; ++i) { s->alive[i] = s->X[i] < eps; ess += s->alive[i]; }
    } else {
        for (int64_t i = 0; i < N; ++i) { s->alive[i] = s->X[i] <= eps; ess += s->alive[i]; }
        flag = 1;
    }
    s->eps = eps;
    s->flag = flag;
    s->n_alive = ess;
    s->resampled = 0;
    /* ref :145-153 */
    if (c->alpha * (double)ess <= (double)N * c->min_r_ess) {
        if (ess == 0) return fail("resampling with zero alive particles");
        int64_t *idxalive = s->ta; /* scratch */
        int64_t n = 0;
        for (int64_t i = 0; i < N; ++i) if (s->alive[i]) idxalive[n++] = i;
        for (int64_t k = 0; k < N; ++k) {
            int64_t src = idxalive[k % n];
            for (int q = 0; q < d; ++q) s->th2[(int64_t)q * N + k] = s->th[(int64_t)q * N + src];
            s->X2[k] = s->X[src];
            s->lpi2[k] = s->lpi[src];
        }
        double *tmp;
        tmp = s->th; s->th = s->th2; s->th2 = tmp;
        tmp = s->X; s->X = s->X2; s->X2 = tmp;
        tmp = s->lpi; s->lpi = s->lpi2; s->lpi2 = tmp;
        memset(s->alive, 1, (size_t)N);
        s->resampled = 1;
    }
    s->accepted = 0; /* ref :156 */
    s->sweeps = 0;
    return 0;
}
void kor_smc_sweep_range(kor_smc_t *s, int64_t lo, int64_t hi, int64_t *acc, int64_t *evals, int64_t *events) {
    smc_sweep(s, s->next_epoch, lo, hi, acc, evals, events);
}
/* returns 1 when the retry loop should stop (ref :192) */
int kor_smc_sweep_commit(kor_smc_t *s, int64_t acc, int64_t evals, int64_t events) {
    s->accepted += acc;
    s->cost_evals += evals;
    s->events += events;
    s->next_epoch += 1;
    s->sweeps += 1;
    return (double)s->accepted >= s->cfg.mcmc_tol * (double)s->N;
}
int kor_smc_finish(kor_smc_t *s, int *stop) {
    const kor_smc_config_t *c = &s->cfg;
    const double eps = s->eps, epsv = s->eps_prev;
    *stop = 0;
    if (s->nlog == s->caplog) {
        s->caplog = s->caplog ? 2 * s->caplog : 64;
        s->log = (kor_smc_log_t *)realloc(s->log, sizeof(kor_smc_log_t) * (size_t)s->caplog);
    }
    kor_smc_log_t *L = &s->log[s->nlog++];
    L->iteration = s->iteration; L->eps = eps; L->n_alive = s->n_alive; L->flag = s->flag; L->resampled = s->resampled;
    L->accepted = s->accepted; L->cost_evals = s->cost_evals; L->sweeps = s->sweeps;
    if (c->verbose) fprintf(stderr, "(iteration, eps, ESS) = (%lld, %.17g, %lld)\n", (long long)s->iteration, eps, (long long)s->n_alive);
    /* ref :194-198 */
    if (2.0 * fabs(epsv - eps) < c->r_epstol * (fabs(epsv) + fabs(eps))) *stop = 1;
    else if (eps <= c->epstol) *stop = 2;
    else if ((double)s->accepted < c->mcmc_tol * (double)s->N) *stop = 3;
    else if (c->max_iterations > 0 && s->iteration >= c->max_iterations) *stop = 4;
    return 0;
}
int kor_smc_iterate(kor_smc_t *s, int *stop) {
    if (kor_smc_cut(s)) return 1;
    for (int64_t r = 0; r < 1 + s->cfg.mcmc_retrys; ++r) {
        int64_t acc, evals, events;
        kor_smc_sweep_range(s, 0, s->N, &acc, &evals, &events);
        if (kor_smc_sweep_commit(s, acc, evals, events)) break;
    }
    return kor_smc_finish(s, stop);
}
int kor_smc_run(kor_smc_t *s) {
    if (kor_smc_init(s)) return 1;
    int stop = 0;
    while (!stop)
        if (kor_smc_iterate(s, &stop)) return 1;
    return 0;
}
void kor_smc_get_state(const kor_smc_t *s, double *th, double *X, double *lpi, uint8_t *alive) {
    if (th) memcpy(th, s->th, sizeof(double) * (size_t)s->N * (size_t)s->d);
    if (X) memcpy(X, s->X, sizeof(double) * (size_t)s->N);
    if (lpi) memcpy(lpi, s->lpi, sizeof(double) * (size_t)s->N);
    if (alive) memcpy(alive, s->alive, (size_t)s->N);
}
void kor_smc_set_state(kor_smc_t *s, const double *th, const double *X, const double *lpi, const uint8_t *alive) {
    if (th) memcpy(s->th, th, sizeof(double) * (size_t)s->N * (size_t)s->d);
    if (X) memcpy(s->X, X, sizeof(double) * (size_t)s->N);
    if (lpi) memcpy(s->lpi, lpi, sizeof(double) * (size_t)s->N);
    if (alive) memcpy(s->alive, alive, (size_t)s->N);
}
void kor_smc_get_scalars(const kor_smc_t *s, double *eps, int32_t *flag, int64_t *iteration, int64_t *n_alive,
                         int64_t *accepted, int64_t *cost_evals, int64_t *next_epoch) {
    if (eps) *eps = s->eps;
    if (flag) *flag = s->flag;
    if (iteration) *iteration = s->iteration;
    if (n_alive) *n_alive = s->n_alive;
    if (accepted) *accepted = s->accepted;
    if (cost_evals) *cost_evals = s->cost_evals;
    if (next_epoch) *next_epoch = s->next_epoch;
}
int64_t kor_smc_get_log(const kor_smc_t *s, kor_smc_log_t *log, int64_t cap) {
    int64_t n = s->nlog < cap ? s->nlog : cap;
    if (log) memcpy(log, s->log, sizeof(kor_smc_log_t) * (size_t)n);
    return s->nlog;
}
void kor_smc_get_trace(const kor_smc_t *s, int64_t *a, int64_t *b, double *z, double *lprob, double *lpi_p,
                       double *xp, uint8_t *decision, double *thp) {
    size_t N = (size_t)s->N;
    if (a) memcpy(a, s->ta, 8 * N);
    if (b) memcpy(b, s->tb, 8 * N);
    if (z) memcpy(z, s->tz, 8 * N);
    if (lprob) memcpy(lprob, s->tlprob, 8 * N);
    if (lpi_p) memcpy(lpi_p, s->tlpip, 8 * N);
    if (xp) memcpy(xp, s->txp, 8 * N);
    if (decision) memcpy(decision, s->tdec, N);
    if (thp) memcpy(thp, s->thp, 8 * N * (size_t)s->d);
}

/* ------------------------------------------------------------------ */
/* AIS -- ref: src/transition.jl, src/types.jl:51-75, src/KissABC.jl    */
/* ------------------------------------------------------------------ */
struct kor_ais {
    uint64_t seed;
    kor_prior_t prior[16];
    int d;
    kor_model_t model;
    kor_ais_config_t cfg;
    int nthreads;
    int64_t N;
    double *th, *lp, *ll;
    /* trace */
    uint8_t *tmove, *tdec;
    int64_t *ta, *tb, *tc;
    double *tcorr, *thp, *tlpp, *tllp, *te;
    int64_t cost_evals, accepted, sweeps, retries;
    int serial;   /* 1: one word stream in the reference's consumption order (see ST_SERIAL) */
    stream_t ser;
    int err;      /* accept() raised, ref src/types.jl:69-70: 1 "ld_correction is invalid", 2 "starting sample invalid." */
};
void kor_ais_set_serial(kor_ais_t *s) {
    s->serial = 1;
    s->nthreads = 1;
    stream_init(&s->ser, s->seed, ST_SERIAL, 0, 0);
}

int kor_ais_create(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model,
                   const kor_ais_config_t *cfg, int nthreads, kor_ais_t **out) {
    if (d < 1 || d > 16) return fail("d out of range");
    /* ref: src/KissABC.jl:43-48 */
    if (cfg->nwalkers < d + 5) {
        snprintf(g_err, sizeof g_err, "nparticles = %lld is insufficient, set number of particles in AIS(.) atleast to %d",
                 (long long)cfg->nwalkers, d + 5);
        return 1;
    }
    kor_ais_t *s = (kor_ais_t *)calloc(1, sizeof *s);
    s->seed = seed;
    memcpy(s->prior, prior, sizeof(kor_prior_t) * (size_t)d);
    s->d = d; s->model = *model; s->cfg = *cfg; s->nthreads = nthreads > 0 ? nthreads : 1;
    size_t N = (size_t)(s->N = cfg->nwalkers);
    s->th = (double *)calloc(N * (size_t)d, 8); s->thp = (double *)calloc(N * (size_t)d, 8);
    s->lp = (double *)calloc(N, 8); s->ll = (double *)calloc(N, 8);
    s->tmove = (uint8_t *)calloc(N, 1); s->tdec = (uint8_t *)calloc(N, 1);
    s->ta = (int64_t *)calloc(N, 8); s->tb = (int64_t *)calloc(N, 8); s->tc = (int64_t *)calloc(N, 8);
    s->tcorr = (double *)calloc(N, 8); s->tlpp = (double *)calloc(N, 8); s->tllp = (double *)calloc(N, 8);
    s->te = (double *)calloc(N, 8);
    *out = s;
    return 0;
}
void kor_ais_destroy(kor_ais_t *s) {
    if (!s) return;
    free(s->th); free(s->thp); free(s->lp); free(s->ll); free(s->tmove); free(s->tdec); free(s->ta);
    free(s->tb); free(s->tc); free(s->tcorr); free(s->tlpp); free(s->tllp); free(s->te); free(s);
}

/* ref: src/types.jl:51-58 -- (logprior, loglikelihood) of the kernelized posterior (cfg.posterior == 0);
 * ref: src/types.jl:84-91 -- (logprior, cost) of the hard-threshold ApproxPosterior (cfg.posterior == 1): the second
 * slot then holds the cost itself (-logprior when the prior is not finite). */
static void ais_loglike(kor_ais_t *s, const double *xraw, uint32_t tag, uint32_t id, uint32_t epoch, double *lp,
                        double *ll, int *evals) {
    double x[16];
    kor_push_p(s->prior, s->d, xraw, x); /* ref src/KissABC.jl:51,56: loglike(model, push_p(model, particle)) */
    double p = kor_prior_logpdf(s->prior, s->d, x);
    double l = s->cfg.posterior == 1 ? -p : p;
    if (isfinite(p)) {
        double c = s->serial ? cost_on_stream(&s->model, &s->ser, x) : cost_dispatch(&s->model, s->seed, tag, s->d, x, id, epoch);
        if (s->cfg.posterior == 1) l = c;
        else {
            double q = c / s->cfg.scale;
            l = -0.5 * (q * q);
        }
        *evals += 1;
    }
    *lp = p;
    *ll = l;
}
/* ref: src/types.jl:60 and :93-94 */
static int ais_valid(const kor_ais_t *s, double lp, double ll) {
    return s->cfg.posterior == 1 ? (isfinite(ll) && isfinite(lp)) : isfinite(lp + ll);
}

/* ref: src/KissABC.jl:50-61.  attempt t of walker i uses epoch t of the PRIOR / COST_INIT streams */
int kor_ais_init(kor_ais_t *s) {
    const int64_t N = s->N;
    const int d = s->d;
    int64_t budget = s->cfg.retry_sampling * N;
    int64_t cap = budget + 1; /* per-walker retries beyond this always exhaust the budget */
    int64_t retries = 0, evals_total = 0;
    int bad = 0;
    s->err = 0;
    if (s->serial) { /* ref src/KissABC.jl:50-61 in its own order: all draws, all log-densities, then the retry loop walker by walker */
        for (int64_t i = 0; i < N; ++i)
            for (int k = 0; k < d; ++k) bad |= prior1_sample(&s->prior[k], &s->ser, &s->th[(int64_t)k * N + i]);
        for (int64_t i = 0; i < N; ++i) {
            double th[16];
            int evals = 0;
            for (int k = 0; k < d; ++k) th[k] = s->th[(int64_t)k * N + i];
            ais_loglike(s, th, ST_COST_INIT, (uint32_t)i, 0, &s->lp[i], &s->ll[i], &evals);
            evals_total += evals;
        }
        for (int64_t i = 0; i < N; ++i)
            while (!ais_valid(s, s->lp[i], s->ll[i]) && retries <= budget) {
                double th[16];
                int evals = 0;
                for (int k = 0; k < d; ++k) { bad |= prior1_sample(&s->prior[k], &s->ser, &th[k]); s->th[(int64_t)k * N + i] = th[k]; }
                ais_loglike(s, th, ST_COST_INIT, (uint32_t)i, 0, &s->lp[i], &s->ll[i], &evals);
                evals_total += evals;
                retries += 1;
            }
    }
#pragma omp parallel for schedule(dynamic, 16) num_threads(s->nthreads) reduction(+ : retries, evals_total) reduction(| : bad)
    for (int64_t i = s->serial ? N : 0; i < N; ++i) {
        double th[16], lp = 0, ll = 0;
        int evals = 0;
        int64_t t = 0;
        for (;; ++t) {
            bad |= kor_prior_sample(s->seed, s->prior, d, (uint32_t)i, (uint32_t)t, th);
            ais_loglike(s, th, ST_COST_INIT, (uint32_t)i, (uint32_t)t, &lp, &ll, &evals);
            if (ais_valid(s, lp, ll) || t >= cap) break;
        }
        retries += t;
        evals_total += evals;
        for (int k = 0; k < d; ++k) s->th[(int64_t)k * N + i] = th[k];
        s->lp[i] = lp;
        s->ll[i] = ll;
    }
    s->retries = retries;
    s->cost_evals = evals_total;
    s->accepted = 0;
    s->sweeps = 0;
    if (bad) return fail("prior sampling failed (truncation too extreme)");
    if (retries > budget)
        return fail("Prior leads to \xe2\x88\x9e costs too often, tune the prior or increase `retry_sampling`.");
    return 0;
}

/* draw an index from [lo, lo+n) */
static int64_t draw_idx(stream_t *st, int64_t lo, int64_t n) { return lo + (int64_t)kor_index(next_u32(st), (uint32_t)n); }

/* ref: src/transition.jl:67-82 (transition!), :61-65 (propose), :51-59, :2-22, :24-43;
 * src/types.jl:62-75 (accept).  The proposal is built from the ensemble `src` (the pre-half-step
 * snapshot for the red/black schedule, the live ensemble for the sequential one). */
static int ais_transition_from(kor_ais_t *s, const double *src, int64_t i, int64_t lo, int64_t n, uint32_t epoch) {
    const int64_t N = s->N;
    const int d = s->d;
    stream_t st0, *stp = &st0;
    if (s->serial) stp = &s->ser;
    else stream_init(&st0, s->seed, ST_PROPOSE, (uint32_t)i, epoch);
#define st (*stp)
    double p[16], xi[16];
    for (int k = 0; k < d; ++k) xi[k] = src[(int64_t)k * N + i];
    double corr = 0.0;
    int64_t a = i, b = i, c = i;
    /* ref :62 rand(rng,(1,1,1,1,2,2,3)) */
    uint32_t slot = kor_index(next_u32(&st), 7);
    int move = slot < 4 ? 1 : (slot < 6 ? 2 : 3);
    if (move == 1) { /* stretch, ref :51-59, a = 3.0 */
        while (a == i) a = draw_idx(&st, lo, n);
        double u = next_uniform(&st);
        double sa = sqrt(3.0), ra = sqrt(1.0 / 3.0);
        double t = u * (sa - ra) + ra;
        double Z = t * t;
        for (int k = 0; k < d; ++k) {
            double xa = src[(int64_t)k * N + a];
            p[k] = xa + (xi[k] - xa) * Z;
        }
        corr = (double)(d - 1) * kor_log(Z);
        b = c = -1;
    } else if (move == 2) { /* DE, ref :2-22 */
        double z0 = next_normal(&st);
        double gam = (2.38 / sqrt((double)(2 * d))) * kor_exp(z0 * 0.1);
        while (a == i) a = draw_idx(&st, lo, n);
        while (b == a || b == i) b = draw_idx(&st, lo, n);
        for (int k = 0; k < d; ++k) {
            double xa = src[(int64_t)k * N + a], xb = src[(int64_t)k * N + b];
            double W = (xa - xb) * gam;
            double S = (fabs(xa - xb) + fabs(xi[k] - xb)) + fabs(xa - xi[k]);
            double T = ((gam * S) / 300.0) * next_normal(&st);
            p[k] = (xi[k] + W) + T;
        }
        c = -1;
    } else { /* walk, ref :24-43 */
        while (a == i) a = draw_idx(&st, lo, n);
        while (b == a || b == i) b = draw_idx(&st, lo, n);
        while (c == b || c == a || c == i) c = draw_idx(&st, lo, n);
        double z1 = next_normal(&st), z2 = next_normal(&st), z3 = next_normal(&st);
        for (int k = 0; k < d; ++k) {
            double xa = src[(int64_t)k * N + a], xb = src[(int64_t)k * N + b], xc = src[(int64_t)k * N + c];
            double xs = (xa + (xb + xc)) / 3.0;
            double W = ((z1 * (xa - xs)) + (z2 * (xb - xs))) + (z3 * (xc - xs));
            p[k] = xi[k] + W;
        }
    }
#undef st
    double lpp, llp;
    int evals = 0;
    ais_loglike(s, p, ST_COST, (uint32_t)i, epoch, &lpp, &llp, &evals);
    /* ref: src/types.jl:69-74.  accept() raises on a non-finite correction or an invalid CURRENT state before it looks at the
     * proposal; the run is then void (kor_ais_sweep / run_* return the error) */
    if (!isfinite(corr)) {
#pragma omp critical
        if (!s->err) s->err = 1;
    } else if (!ais_valid(s, s->lp[i], s->ll[i])) {
#pragma omp critical
        if (!s->err) s->err = 2;
    }
    int dec;
    double e = NAN;
    if (!ais_valid(s, lpp, llp)) dec = 0;
    else {
        stream_t sa0, *sap = &sa0;
        if (s->serial) sap = &s->ser;
        else stream_init(&sa0, s->seed, ST_ACCEPT, (uint32_t)i, epoch);
#define sa (*sap)
        e = next_exp(&sa);
#undef sa
        if (s->cfg.posterior == 1) { /* ref: src/types.jl:101-103 */
            double lW = (corr + lpp) - s->lp[i];
            double lW2 = fmax(s->cfg.scale, s->ll[i]) - llp;
            dec = ((-e <= lW) && lW2 >= 0) ? 2 : 1;
        } else {                     /* ref: src/types.jl:73-74 */
            double lW = (corr + (lpp + llp)) - (s->lp[i] + s->ll[i]);
            dec = (-e <= lW) ? 2 : 1;
        }
    }
    s->tmove[i] = (uint8_t)move; s->ta[i] = a; s->tb[i] = b; s->tc[i] = c; s->tcorr[i] = corr;
    for (int k = 0; k < d; ++k) s->thp[(int64_t)k * N + i] = p[k];
    s->tlpp[i] = lpp; s->tllp[i] = llp; s->te[i] = e; s->tdec[i] = (uint8_t)dec;
#pragma omp atomic
    s->cost_evals += evals;
    if (dec == 2) {
        for (int k = 0; k < d; ++k) s->th[(int64_t)k * N + i] = p[k];
        s->lp[i] = lpp;
        s->ll[i] = llp;
#pragma omp atomic
        s->accepted += 1;
        return 1;
    }
    return 0;
}
int kor_ais_transition(kor_ais_t *s, int64_t i, int64_t lo, int64_t n, uint32_t epoch) {
    return ais_transition_from(s, s->th, i, lo, n, epoch);
}

/* red/black sweep: colour 0 = walkers [0,h) move against [h,N); colour 1 = the converse; h = N/2.
 * Within a half-step the moving walkers only READ the other colour, so updates are independent. */
int kor_ais_sweep(kor_ais_t *s) {
    const int64_t N = s->N, h = N / 2;
    if (N - h < 3 || h < 3) return fail("red/black AIS needs >= 3 walkers per colour");
    uint32_t e0 = (uint32_t)(2 * s->sweeps);
#pragma omp parallel for schedule(dynamic, 16) num_threads(s->nthreads)
    for (int64_t i = 0; i < h; ++i) ais_transition_from(s, s->th, i, h, N - h, e0);
#pragma omp parallel for schedule(dynamic, 16) num_threads(s->nthreads)
    for (int64_t i = h; i < N; ++i) ais_transition_from(s, s->th, i, 0, h, e0 + 1);
    s->sweeps += 1;
    if (s->err) return fail(s->err == 1 ? "ld_correction is invalid" : "starting sample invalid.");
    return 0;
}

/* number of `step` calls AbstractMCMC.mcmcsample makes before saved sample m (0-based):
 * discard_initial + m*thinning  [dep: AbstractMCMC 2.1-3.1, restated] */
static int64_t steps_before(const kor_ais_config_t *c, int64_t m) { return c->discard_initial + m * c->thinning; }

/* reference schedule, ref: src/KissABC.jl:66-80: step s (1-based) applies ntransitions moves to walker
 * (s-1) mod N against the whole live ensemble, emits it, rotates. */
int kor_ais_run_sequential(kor_ais_t *s, double *out) {
    const int64_t N = s->N, Ns = s->cfg.nsamples;
    const int d = s->d;
    if (kor_ais_init(s)) return 1;
    int64_t step = 0; /* steps done */
    uint32_t epoch = 0;
    for (int64_t m = 0; m < Ns; ++m) {
        int64_t target = steps_before(&s->cfg, m);
        int64_t w = N - 1; /* ref :63 first sample is particles[end] */
        while (step < target) {
            w = step % N;
            for (int64_t r = 0; r < s->cfg.ntransitions; ++r) ais_transition_from(s, s->th, w, 0, N, epoch++);
            ++step;
        }
        if (target > 0) w = (target - 1) % N;
        for (int k = 0; k < d; ++k) { /* ref src/KissABC.jl:78 records push_p(model, sample[i]) */
            double v = s->th[(int64_t)k * N + w];
            out[(int64_t)k * Ns + m] = prior_is_discrete(&s->prior[k]) ? nearbyint(v) : v;
        }
    }
    return 0;
}

/* device schedule: steps are grouped in rounds of N; a round = ntransitions red/black sweeps of the whole
 * ensemble; step s emits walker (s-1) mod N as it stands at the end of round ceil(s/N). */
int kor_ais_run_parallel(kor_ais_t *s, double *out) {
    const int64_t N = s->N, Ns = s->cfg.nsamples;
    const int d = s->d;
    if (kor_ais_init(s)) return 1;
    int64_t rounds = 0;
    for (int64_t m = 0; m < Ns; ++m) {
        int64_t target = steps_before(&s->cfg, m);
        int64_t w = N - 1, need = 0;
        if (target > 0) { w = (target - 1) % N; need = (target + N - 1) / N; }
        while (rounds < need) {
            for (int64_t r = 0; r < s->cfg.ntransitions; ++r)
                if (kor_ais_sweep(s)) return 1;
            ++rounds;
        }
        for (int k = 0; k < d; ++k) { /* ref src/KissABC.jl:78 records push_p(model, sample[i]) */
            double v = s->th[(int64_t)k * N + w];
            out[(int64_t)k * Ns + m] = prior_is_discrete(&s->prior[k]) ? nearbyint(v) : v;
        }
    }
    return 0;
}
void kor_ais_get_state(const kor_ais_t *s, double *th, double *lp, double *ll) {
    if (th) memcpy(th, s->th, 8 * (size_t)s->N * (size_t)s->d);
    if (lp) memcpy(lp, s->lp, 8 * (size_t)s->N);
    if (ll) memcpy(ll, s->ll, 8 * (size_t)s->N);
}
void kor_ais_set_state(kor_ais_t *s, const double *th, const double *lp, const double *ll) {
    if (th) memcpy(s->th, th, 8 * (size_t)s->N * (size_t)s->d);
    if (lp) memcpy(s->lp, lp, 8 * (size_t)s->N);
    if (ll) memcpy(s->ll, ll, 8 * (size_t)s->N);
}
void kor_ais_get_counters(const kor_ais_t *s, int64_t *cost_evals, int64_t *accepted, int64_t *sweeps, int64_t *retries) {
    if (cost_evals) *cost_evals = s->cost_evals;
    if (accepted) *accepted = s->accepted;
    if (sweeps) *sweeps = s->sweeps;
    if (retries) *retries = s->retries;
}
void kor_ais_get_trace(const kor_ais_t *s, uint8_t *move, int64_t *a, int64_t *b, int64_t *c, double *corr,
                       double *thp, double *lp_p, double *ll_p, double *e, uint8_t *decision) {
    size_t N = (size_t)s->N;
    if (move) memcpy(move, s->tmove, N);
    if (a) memcpy(a, s->ta, 8 * N);
    if (b) memcpy(b, s->tb, 8 * N);
    if (c) memcpy(c, s->tc, 8 * N);
    if (corr) memcpy(corr, s->tcorr, 8 * N);
    if (thp) memcpy(thp, s->thp, 8 * N * (size_t)s->d);
    if (lp_p) memcpy(lp_p, s->tlpp, 8 * N);
    if (ll_p) memcpy(ll_p, s->tllp, 8 * N);
    if (e) memcpy(e, s->te, 8 * N);
    if (decision) memcpy(decision, s->tdec, N);
}

/* ------------------------------------------------------------------ */
/* ABCDE and pfilter -- ref: src/smc.jl:275-428                         */
/* Shared spec: init draws particle i from stream (PRIOR, i, t) and      */
/* costs it with (COST_INIT, i, t), t = 0,1,... until cost and logprior  */
/* are finite (ref :284-297, :358-371).  The cost sees the RAW particle  */
/* (`cost(p.x)`), the prior the push_p'ed one, as in the reference.      */
/* ------------------------------------------------------------------ */
#define PMC_INIT_TRIES 1000
static int pmc_init(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model, int64_t N, int nthreads,
                    double *th, double *lp, double *C, int64_t *evals_out) {
    int bad = 0;
    int64_t evals = 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : bad) reduction(+ : evals)
    for (int64_t i = 0; i < N; ++i) {
        double x[16], xp[16];
        int ok = 0;
        for (int t = 0; t < PMC_INIT_TRIES && !ok; ++t) {
            if (kor_prior_sample(seed, prior, d, (uint32_t)i, (uint32_t)t, x)) { bad = 1; break; }
            kor_push_p(prior, d, x, xp);
            double l = kor_prior_logpdf(prior, d, xp), c = NAN;
            if (isfinite(l)) {
                c = cost_dispatch(model, seed, ST_COST_INIT, d, x, (uint32_t)i, (uint32_t)t);
                evals += 1;
            }
            if (isfinite(l) && isfinite(c)) {
                for (int k = 0; k < d; ++k) th[(int64_t)k * N + i] = x[k];
                lp[i] = l;
                C[i] = c;
                ok = 1;
            }
        }
        if (!ok) bad = 1;
    }
    *evals_out = evals;
    return bad;
}

typedef struct { uint64_t key; int64_t idx; } pmc_kv_t;
static uint64_t dkey_of(double x) {
    uint64_t b = d2u(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static int pmc_kv_cmp(const void *a, const void *b) {
    const pmc_kv_t *p = (const pmc_kv_t *)a, *q = (const pmc_kv_t *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx ? 1 : 0);
}

int kor_abcde_run(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model, const kor_abcde_config_t *cfg,
                  int nthreads, double *theta_out, double *cost_out, int32_t *reached, int64_t *nsim, int64_t *generations_done) {
    const int64_t N = cfg->nparticles;
    if (!(cfg->alpha >= 0.0 && cfg->alpha < 1.0)) return fail("\u03b1 must be in 0 <= \u03b1 < 1."); /* ref :353 */
    if (N < 3) return fail("ABCDE needs at least 3 particles");
    double *th = malloc(sizeof(double) * (size_t)N * d), *nth = malloc(sizeof(double) * (size_t)N * d);
    double *lp = malloc(sizeof(double) * N), *nlp = malloc(sizeof(double) * N);
    double *D = malloc(sizeof(double) * N), *nD = malloc(sizeof(double) * N);
    pmc_kv_t *kv = malloc(sizeof(pmc_kv_t) * N);
    int64_t sims = 0, gen = 0;
    int rc = pmc_init(seed, prior, d, model, N, nthreads, th, lp, D, &sims);
    if (rc) fail("Prior leads to non-finite costs too often");
    sims = 0; /* ref :373 nsims counts the simulations of the generations only */
    const double gam = (cfg->proposal_width * 2.38) / sqrt((double)(2 * d)); /* ref :374 */
    while (!rc && gen < cfg->generations) {
        gen += 1;
        memcpy(nth, th, sizeof(double) * (size_t)N * d);
        memcpy(nlp, lp, sizeof(double) * N);
        memcpy(nD, D, sizeof(double) * N);
        for (int64_t i = 0; i < N; ++i) { kv[i].key = dkey_of(D[i]); kv[i].idx = i; }
        qsort(kv, (size_t)N, sizeof(pmc_kv_t), pmc_kv_cmp);
        const double eps_l = D[kv[0].idx], eps_h = D[kv[N - 1].idx]; /* ref :382 extrema */
        if (cfg->earlystop && eps_h <= cfg->eps_target) { gen -= 1; break; }
        const double eps_pop = fmax(cfg->eps_target, eps_l + cfg->alpha * (eps_h - eps_l));
        int64_t gsims = 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : gsims)
        for (int64_t i = 0; i < N; ++i) {
            if (cfg->earlystop && D[i] <= cfg->eps_target) continue;
            stream_t st;
            stream_init(&st, seed, ST_PROPOSE, (uint32_t)i, (uint32_t)gen);
            int64_t s = i;
            const double eps = D[i] <= cfg->eps_target ? cfg->eps_target : eps_pop;
            if (D[i] > eps) {
                /* ref :393 rand((1:n)[D .<= D[i]]): uniform over that set; spec: the set in (cost, index) order */
                const uint64_t ki = dkey_of(D[i]);
                int64_t lo = 0, hi = N; /* first position with key > ki */
                while (lo < hi) { int64_t mid = (lo + hi) / 2; if (kv[mid].key <= ki) lo = mid + 1; else hi = mid; }
                s = kv[kor_index(next_u32(&st), (uint32_t)lo)].idx;
            }
            int64_t a = s, b;
            while (a == s) a = (int64_t)kor_index(next_u32(&st), (uint32_t)N);
            b = a;
            while (b == a || b == s) b = (int64_t)kor_index(next_u32(&st), (uint32_t)N);
            double thp[16], thq[16];
            for (int k = 0; k < d; ++k)
                thp[k] = th[(int64_t)k * N + s] + (th[(int64_t)k * N + a] - th[(int64_t)k * N + b]) * gam;
            kor_push_p(prior, d, thp, thq);
            const double l = kor_prior_logpdf(prior, d, thq);
            const double w = l - lp[i];
            if (kor_log(next_uniform(&st)) > fmin(0.0, w)) continue;
            gsims += 1;
            const double dp = cost_dispatch(model, seed, ST_COST, d, thp, (uint32_t)i, (uint32_t)gen);
            if (dp <= fmax(eps, D[i])) {
                nD[i] = dp;
                for (int k = 0; k < d; ++k) nth[(int64_t)k * N + i] = thp[k];
                nlp[i] = l;
            }
        }
        sims += gsims;
        double *t;
        t = th; th = nth; nth = t;
        t = lp; lp = nlp; nlp = t;
        t = D; D = nD; nD = t;
    }
    if (!rc) {
        double mx = -INFINITY;
        for (int64_t i = 0; i < N; ++i) {
            double x[16], xp[16];
            for (int k = 0; k < d; ++k) x[k] = th[(int64_t)k * N + i];
            kor_push_p(prior, d, x, xp);
            for (int k = 0; k < d; ++k) theta_out[(int64_t)k * N + i] = xp[k];
            cost_out[i] = D[i];
            if (D[i] > mx) mx = D[i];
        }
        if (reached) *reached = mx <= cfg->eps_target;
        if (nsim) *nsim = sims;
        if (generations_done) *generations_done = gen;
    }
    free(th); free(nth); free(lp); free(nlp); free(D); free(nD); free(kv);
    return rc;
}

int64_t kor_pfilter_nparticles(int64_t n, int d, double q) {
    const int64_t lowN = 4 * (int64_t)d;
    if ((double)n * q <= (double)lowN) n = (int64_t)ceil((double)(lowN + 1) / q);
    return n;
}

#define PF_MAX_ROUNDS 100000
int kor_pfilter_run(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model, const kor_pfilter_config_t *cfg,
                    int nthreads, double *theta_out, double *cost_out, double *eps_out, int64_t *iters_out, int64_t *nreps_out,
                    int64_t *cost_evals) {
    if (!(cfg->q > 0.0 && cfg->q <= 1.0)) return fail("pfilter needs 0 < q <= 1");
    const int64_t N = kor_pfilter_nparticles(cfg->nparticles, d, cfg->q);
    double *th = malloc(sizeof(double) * (size_t)N * d), *lp = malloc(sizeof(double) * N), *C = malloc(sizeof(double) * N);
    int64_t *idxok = malloc(sizeof(int64_t) * N), *pend = malloc(sizeof(int64_t) * N), *pend2 = malloc(sizeof(int64_t) * N);
    pmc_kv_t *kv = malloc(sizeof(pmc_kv_t) * N);
    double *srt = malloc(sizeof(double) * N);
    int64_t evals = 0, iters = 0, reps_total = 0;
    uint32_t round = 0; /* epoch of the attempt streams: one per rejection round over the whole run */
    double eps = INFINITY;
    int rc = pmc_init(seed, prior, d, model, N, nthreads, th, lp, C, &evals);
    if (rc) fail("Prior leads to non-finite costs too often");
    while (!rc) {
        iters += 1;
        /* ref :300-302; spec: `rand(trng, idxok)` is taken over idxok in (cost, index) order, i.e. the head of one
         * sort of the costs, whose tail is idxbad */
        for (int64_t i = 0; i < N; ++i) { kv[i].key = dkey_of(C[i]); kv[i].idx = i; }
        qsort(kv, (size_t)N, sizeof(pmc_kv_t), pmc_kv_cmp);
        for (int64_t i = 0; i < N; ++i) srt[i] = C[kv[i].idx];
        eps = quantile7_sorted(srt, N, cfg->q);
        int64_t n_ok = 0, n_bad = 0;
        for (int64_t i = 0; i < N; ++i) {
            if (C[kv[i].idx] > eps) pend[n_bad++] = kv[i].idx;
            else idxok[n_ok++] = kv[i].idx;
        }
        if (n_bad > 0 && n_ok < 3) { rc = fail("pfilter: fewer than 3 particles under the quantile"); break; }
        int64_t nreps = 0, n_pend = n_bad;
        for (int r = 0; n_pend > 0; ++r) {
            if (r >= PF_MAX_ROUNDS) { rc = fail("pfilter: rejection loop does not terminate"); break; }
            round += 1;
            nreps += n_pend; /* ref :313 localreps: every attempt counts */
            int64_t gev = 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : gev)
            for (int64_t w = 0; w < n_pend; ++w) {
                const int64_t i = pend[w];
                stream_t st;
                stream_init(&st, seed, ST_PROPOSE, (uint32_t)i, round);
                int64_t b = (int64_t)kor_index(next_u32(&st), (uint32_t)n_ok), c = b, e;
                while (c == b) c = (int64_t)kor_index(next_u32(&st), (uint32_t)n_ok);
                e = b;
                while (e == b || e == c) e = (int64_t)kor_index(next_u32(&st), (uint32_t)n_ok);
                b = idxok[b]; c = idxok[c]; e = idxok[e];
                const double sc = next_normal(&st) * cfg->proposal_width; /* ref :312 */
                double p[16], pq[16];
                for (int k = 0; k < d; ++k)
                    p[k] = th[(int64_t)k * N + b] + (th[(int64_t)k * N + e] - th[(int64_t)k * N + c]) * sc;
                kor_push_p(prior, d, p, pq);
                const double ll = kor_prior_logpdf(prior, d, pq);
                pend2[w] = i; /* still pending unless accepted below */
                if (kor_log(next_uniform(&st)) > fmin(0.0, ll - lp[i])) continue;
                const double Cp = cost_dispatch(model, seed, ST_COST, d, p, (uint32_t)i, round);
                gev += 1;
                if (Cp > eps) continue; /* ref :320: a NaN cost compares false and is accepted */
                C[i] = Cp;
                for (int k = 0; k < d; ++k) th[(int64_t)k * N + i] = p[k];
                lp[i] = ll;
                pend2[w] = -1;
            }
            evals += gev;
            int64_t m = 0;
            for (int64_t w = 0; w < n_pend; ++w)
                if (pend2[w] >= 0) pend[m++] = pend2[w];
            n_pend = m;
        }
        if (rc) break;
        reps_total += nreps;
        /* ref :331-335; an iteration without bad particles (eff = 0/0 in the reference, which then never leaves the
         * loop) ends the run here */
        if (n_bad == 0) break;
        const double eff = (double)n_bad / (double)nreps;
        if (eff < cfg->eff_tol) break;
        if (eps < cfg->epstol) break;
        if (cfg->max_iters > 0 && iters > cfg->max_iters) break;
    }
    if (!rc) {
        for (int64_t i = 0; i < N; ++i) {
            double x[16], xp[16];
            for (int k = 0; k < d; ++k) x[k] = th[(int64_t)k * N + i];
            kor_push_p(prior, d, x, xp);
            for (int k = 0; k < d; ++k) theta_out[(int64_t)k * N + i] = xp[k];
            cost_out[i] = C[i];
        }
        if (eps_out) *eps_out = eps;
        if (iters_out) *iters_out = iters;
        if (nreps_out) *nreps_out = reps_total;
        if (cost_evals) *cost_evals = evals;
    }
    free(th); free(lp); free(C); free(idxok); free(pend); free(pend2); free(kv); free(srt);
    return rc;
}
