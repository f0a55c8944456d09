// kabc_models.cuh -- registered device simulators + distances: the replacement of the opaque `cost`
// closure of the reference (src/types.jl:55, src/smc.jl:123,176).  One thread evaluates one particle
// (normal, MA(2), Lotka-Volterra); g-and-k uses one warp per particle (kabc_gk.cuh).
//
// Each simulator exists in two precisions:
//   KABC_F64        fixed sequence of IEEE double ops == oracle/kabc_oracle.c bit for bit
//   KABC_F32_ACC64  draws generated in FP32 on the MUFU pipe from the SAME Philox words, accumulated in FP32
//                   partial sums around a known shift, distance finished in FP64
#pragma once
#include "kabc_device.cuh"

namespace kabc {

struct DModel {
    int kind, precision, n_draws, n_target;
    uint32_t push_mask; // bit k: component k follows a discrete law and reaches the cost rounded (push_p, src/types.jl:32);
                        // set by the smc / ais handles from their prior, 0 for the bare kabc_eval_cost entry point
    double target[KABC_MAX_TARGET];
    double param[KABC_MAX_PARAM];
};

// ------------------------------------------------------------------ normal model, ref README.md:35-52
// x = randn(n).*sigma .+ mu ; hypot(mean(x)-t0, (std(x)-t1)*w) ; std with n-1.
__device__ __forceinline__ double cost_normal_f64(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                                  uint32_t epoch, double mu, double sigma) {
    const int n = m.n_draws;
    const int nb = (n + 3) >> 2;
    double sum = 0.0;
    for (int b = 0; b < nb; ++b) {
        uint32_t w0, w1, w2, w3;
        philox4x32_10(rk, (uint32_t)b, id, epoch, tag, w0, w1, w2, w3);
        double z[4];
        normal_pair64(w0, w1, z[0], z[1]);
        normal_pair64(w2, w3, z[2], z[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (4 * b + q < n) sum = xadd(sum, xadd(xmul(z[q], sigma), mu));
    }
    const double mean = xdiv(sum, (double)n);
    double ss = 0.0;
    for (int b = 0; b < nb; ++b) {
        uint32_t w0, w1, w2, w3;
        philox4x32_10(rk, (uint32_t)b, id, epoch, tag, w0, w1, w2, w3);
        double z[4];
        normal_pair64(w0, w1, z[0], z[1]);
        normal_pair64(w2, w3, z[2], z[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (4 * b + q < n) {
                double dx = xsub(xadd(xmul(z[q], sigma), mu), mean);
                ss = xadd(ss, xmul(dx, dx));
            }
    }
    const double sd = xsqrt(xdiv(ss, (double)(n - 1)));
    const double d1 = xsub(mean, m.target[0]);
    const double d2 = xmul(xsub(sd, m.target[1]), m.param[0]);
    return xsqrt(xadd(xmul(d1, d1), xmul(d2, d2)));
}

// FP32 draws.  mean(x) = mu + sigma*mean(z) and std(x) = sigma*std(z) exactly (x = mu + sigma*z), so the kernel
// accumulates the one-pass sums of z and z^2 in FP32 (well conditioned: z is centred, unit scale) and applies mu and
// sigma once, in FP64, at the end.  Full Philox blocks go through the pair-sum identities of normal_pair_sums32.
__device__ __forceinline__ double cost_normal_f32(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                                  uint32_t epoch, double mu, double sigma) {
    const int n = m.n_draws;
    const int nbf = n >> 2; // full blocks
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float ps[2] = {0.f, 0.f}, pl[2] = {0.f, 0.f}; // full blocks: pair sums of s sin(t + pi/4) and of lg2(u1)
#ifndef KABC_NORMAL_UNROLL
#define KABC_NORMAL_UNROLL 4
#endif
#define KABC_STR2(x) #x
#define KABC_STR(x) KABC_STR2(x)
    _Pragma(KABC_STR(unroll KABC_NORMAL_UNROLL))
    for (int b = 0; b < nbf; ++b) {
        uint32_t w0, w1, w2, w3;
        philox4x32_10(rk, (uint32_t)b, id, epoch, tag, w0, w1, w2, w3);
        normal_pair_sums32(w0, w1, ps[0], pl[0]);
        normal_pair_sums32(w2, w3, ps[1], pl[1]);
    }
    if (n & 3) {
        uint32_t w0, w1, w2, w3;
        philox4x32_10(rk, (uint32_t)nbf, id, epoch, tag, w0, w1, w2, w3);
        float z[4];
        normal_pair32(w0, w1, z[0], z[1]);
        normal_pair32(w2, w3, z[2], z[3]);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q < (n & 3)) {
                s1[q] = __fadd_rn(s1[q], z[q]);
                s2[q] = __fmaf_rn(z[q], z[q], s2[q]);
            }
    }
    // sum z = 2 sqrt(ln 2) * sum(s sin),  sum z^2 = -2 ln 2 * sum(lg2 u1)   (+ the draws of the partial last block)
    const double S1 = 1.6651092223153954 * ((double)ps[0] + (double)ps[1]) +
                      (((double)s1[0] + (double)s1[1]) + ((double)s1[2] + (double)s1[3]));
    const double S2 = -1.3862943611198906 * ((double)pl[0] + (double)pl[1]) +
                      (((double)s2[0] + (double)s2[1]) + ((double)s2[2] + (double)s2[3]));
    const double dn = (double)n;
    const double mean = mu + sigma * (S1 / dn);
    double var = (S2 - S1 * S1 / dn) / (double)(n - 1);
    if (var < 0.0) var = 0.0;
    const double sd = fabs(sigma) * sqrt(var);
    const double d1 = mean - m.target[0];
    const double d2 = (sd - m.target[1]) * m.param[0];
    return sqrt(d1 * d1 + d2 * d2);
}

// ------------------------------------------------------------------ MA(2), SURVEY.md Appendix B
// y_t = e_{t+2} + th1 e_{t+1} + th2 e_t ; tau_j = (1/n) sum_{t>=j} y_t y_{t-j} ; || tau - target ||_2
__device__ __forceinline__ bool ma2_in_triangle(double t1, double t2) {
    return t1 > -2.0 && t1 < 2.0 && xadd(t1, t2) > -1.0 && xsub(t1, t2) < 1.0;
}
__device__ __forceinline__ double cost_ma2_f64(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                               uint32_t epoch, double t1, double t2) {
    if (!ma2_in_triangle(t1, t2)) return dinf();
    const int n = m.n_draws;
    const int nb = (n + 2 + 3) >> 2;
    double e0 = 0, e1 = 0, y1 = 0, y2 = 0, a1 = 0, a2 = 0;
    int t = -2;
    for (int b = 0; b < nb; ++b) {
        uint32_t w0, w1, w2, w3;
        philox4x32_10(rk, (uint32_t)b, id, epoch, tag, w0, w1, w2, w3);
        double z[4];
        normal_pair64(w0, w1, z[0], z[1]);
        normal_pair64(w2, w3, z[2], z[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (4 * b + q < n + 2) {
                double e2 = z[q];
                if (t >= 0) {
                    double y = xadd(xadd(e2, xmul(t1, e1)), xmul(t2, e0));
                    if (t >= 1) a1 = xadd(a1, xmul(y, y1));
                    if (t >= 2) a2 = xadd(a2, xmul(y, y2));
                    y2 = y1;
                    y1 = y;
                }
                e0 = e1;
                e1 = e2;
                ++t;
            }
        }
    }
    const double d1 = xsub(xdiv(a1, (double)n), m.target[0]);
    const double d2 = xsub(xdiv(a2, (double)n), m.target[1]);
    return xsqrt(xadd(xmul(d1, d1), xmul(d2, d2)));
}
__device__ __forceinline__ double cost_ma2_f32(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                               uint32_t epoch, double t1d, double t2d) {
    if (!ma2_in_triangle(t1d, t2d)) return dinf();
    const int n = m.n_draws;
    const int nb = (n + 2 + 3) >> 2;
    const float t1 = (float)t1d, t2 = (float)t2d;
    float e0 = 0.f, e1 = 0.f, y1 = 0.f, y2 = 0.f, a1 = 0.f, a2 = 0.f;
    int t = -2;
    for (int b = 0; b < nb; ++b) {
        uint32_t w0, w1, w2, w3;
        philox4x32_10(rk, (uint32_t)b, id, epoch, tag, w0, w1, w2, w3);
        float z[4];
        normal_pair32(w0, w1, z[0], z[1]);
        normal_pair32(w2, w3, z[2], z[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (4 * b + q < n + 2) {
                float e2 = z[q];
                if (t >= 0) {
                    float y = __fmaf_rn(t2, e0, __fmaf_rn(t1, e1, e2));
                    if (t >= 1) a1 = __fmaf_rn(y, y1, a1);
                    if (t >= 2) a2 = __fmaf_rn(y, y2, a2);
                    y2 = y1;
                    y1 = y;
                }
                e0 = e1;
                e1 = e2;
                ++t;
            }
        }
    }
    const double d1 = (double)a1 / (double)n - m.target[0];
    const double d2 = (double)a2 / (double)n - m.target[1];
    return sqrt(d1 * d1 + d2 * d2);
}

// ------------------------------------------------------------------ Lotka-Volterra, Gillespie direct method
// theta = log rates; param = {X0, Y0, T, G, max_events}; target = [X(t_1..t_G), Y(t_1..t_G)], t_g = g T/G.
// The simulator is a resumable state machine (one SSA event per step()) so that the sweep kernel can re-fill a lane
// with a new particle as soon as its trajectory ends (event counts differ by orders of magnitude across the prior).
template <bool F32>
struct LvSim {
    double c1, c2, c3, X, Y, t, acc, dt;
    long long ev, max_events;
    uint32_t blk, w2, w3, id, epoch, tag;
    int g, G, have;
    bool capped;

    __device__ __forceinline__ void init(const DModel &m, uint32_t tag_, uint32_t id_, uint32_t epoch_, double l1, double l2,
                                         double l3) {
        c1 = xexp(l1); c2 = xexp(l2); c3 = xexp(l3);
        X = m.param[0]; Y = m.param[1];
        G = (int)m.param[3];
        max_events = (long long)m.param[4];
        dt = xdiv(m.param[2], (double)G);
        t = 0.0; acc = 0.0; g = 0; ev = 0; blk = 0; w2 = 0; w3 = 0; have = 0; capped = false;
        id = id_; epoch = epoch_; tag = tag_;
    }
    // one Gillespie event (or the final recording of the grid); returns true when the trajectory is finished
    __device__ __forceinline__ bool step(const DModel &m, const RoundKeys &rk) {
        if (g >= G) return true;
        const double a1 = xmul(c1, X), a2 = xmul(xmul(c2, X), Y), a3 = xmul(c3, Y);
        const double a0 = xadd(xadd(a1, a2), a3);
        double tn;
        uint32_t u0 = 0, u1 = 0;
        if (a0 > 0.0) {
            if (ev >= max_events) { capped = true; return true; }
            if (have == 0) {
                uint32_t w0, w1;
                philox4x32_10(rk, blk, id, epoch, tag, w0, w1, w2, w3);
                blk += 1;
                have = 1;
                u0 = w0; u1 = w1;
            } else {
                have = 0;
                u0 = w2; u1 = w3;
            }
            double lg;
            if (F32) {
                float uf = __fmaf_rn(__uint2float_rn(u0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                lg = (double)__fmul_rn(mufu_lg2(uf), 0.6931471805599453f);
            } else {
                lg = xlog(u01(u0));
            }
            tn = xadd(t, xdiv(-lg, a0));
        } else {
            tn = dinf();
        }
        while (g < G && xmul((double)(g + 1), dt) <= tn) { // record the pre-event state
            const double dx = xsub(X, m.target[g]), dy = xsub(Y, m.target[G + g]);
            acc = xadd(acc, xmul(dx, dx));
            acc = xadd(acc, xmul(dy, dy));
            ++g;
        }
        if (g >= G) return true;
        const double r = xmul(u01(u1), a0);
        if (r < a1) X = xadd(X, 1.0);
        else if (r < xadd(a1, a2)) { X = xsub(X, 1.0); Y = xadd(Y, 1.0); }
        else Y = xsub(Y, 1.0);
        t = tn;
        ++ev;
        return false;
    }
    __device__ __forceinline__ double result() const { return capped ? dinf() : xsqrt(xdiv(acc, (double)(2 * G))); }
};

// F32 instantiation: the copy numbers are exact in FP32 (< 2^24: the event cap bounds them), propensities, waiting time
// and reaction choice are FP32 with MUFU lg2 / rcp (no FP64 division per event), the clock is FP64 (one DADD per event: a
// 20 000-event trajectory must not drift across the 16 recording times) and so are the <= 32 squared grid differences.
// Same Philox words in the same roles as the F64 machine; trajectories decorrelate from it after the first waiting time
// that rounds differently, so this mode is compared with F64 in distribution (tests/test_gpu_parity.py).
template <>
struct LvSim<true> {
    float c1, c2, c3, X, Y;
    double t, dt, acc;
    long long ev, max_events;
    uint32_t blk, w2, w3, id, epoch, tag;
    int g, G, have;
    bool capped;

    __device__ __forceinline__ void init(const DModel &m, uint32_t tag_, uint32_t id_, uint32_t epoch_, double l1, double l2,
                                         double l3) {
        c1 = (float)xexp(l1); c2 = (float)xexp(l2); c3 = (float)xexp(l3);
        X = (float)m.param[0]; Y = (float)m.param[1];
        G = (int)m.param[3];
        max_events = (long long)m.param[4];
        dt = xdiv(m.param[2], (double)G);
        t = 0.0; acc = 0.0; g = 0; ev = 0; blk = 0; w2 = 0; w3 = 0; have = 0; capped = false;
        id = id_; epoch = epoch_; tag = tag_;
    }
    __device__ __forceinline__ bool step(const DModel &m, const RoundKeys &rk) {
        if (g >= G) return true;
        const float a1 = __fmul_rn(c1, X), a2 = __fmul_rn(__fmul_rn(c2, X), Y), a3 = __fmul_rn(c3, Y);
        const float a0 = __fadd_rn(__fadd_rn(a1, a2), a3);
        double tn;
        uint32_t u1 = 0;
        if (a0 > 0.f) {
            if (ev >= max_events) { capped = true; return true; }
            uint32_t u0;
            if (have == 0) {
                uint32_t w0, w1;
                philox4x32_10(rk, blk, id, epoch, tag, w0, w1, w2, w3);
                blk += 1;
                have = 1;
                u0 = w0; u1 = w1;
            } else {
                have = 0;
                u0 = w2; u1 = w3;
            }
            const float uf = __fmaf_rn(__uint2float_rn(u0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
            float rcp;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(a0));
            const float wait = __fmul_rn(__fmul_rn(mufu_lg2(uf), -0.6931471805599453f), rcp);
            tn = t + (double)wait;
        } else {
            tn = dinf();
        }
        while (g < G && (double)(g + 1) * dt <= tn) { // record the pre-event state
            const double dx = (double)X - m.target[g], dy = (double)Y - m.target[G + g];
            acc += dx * dx;
            acc += dy * dy;
            ++g;
        }
        if (g >= G) return true;
        // 24-bit uniform strictly inside (0,1): (w+0.5) 2^-32 rounds to 1.0f for the top 128 words, and r = a0 would fire a
        // reaction whose propensity is zero
        const float r = __fmul_rn(__fmaf_rn((float)(u1 >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f), a0);
        if (r < a1) X += 1.f;
        else if (r < __fadd_rn(a1, a2)) { X -= 1.f; Y += 1.f; }
        else Y -= 1.f;
        t = tn;
        ++ev;
        return false;
    }
    __device__ __forceinline__ double result() const { return capped ? dinf() : sqrt(acc / (double)(2 * G)); }
};

template <bool F32>
__device__ __forceinline__ double cost_lv(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                          uint32_t epoch, double l1, double l2, double l3, long long &events) {
    LvSim<F32> sim;
    sim.init(m, tag, id, epoch, l1, l2, l3);
    while (!sim.step(m, rk)) {}
    events = sim.ev;
    return sim.result();
}

// ------------------------------------------------------------------ deterministic costs of the reference's tests
template <typename F>
__device__ __forceinline__ double cost_det(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id, uint32_t epoch,
                                           F th) {
    const double th0 = th(0);
    if (m.param[0] == 0.0) return fabs(xsub(xadd(xmul(th0, th0), 1.0), m.target[0]));
    if (m.param[0] == 2.0) { // test/runtests.jl:105-112: (n*n+du)*(n+randn()*0.01)
        Stream st(rk, tag, id, epoch);
        double noisy = xadd(th0, xmul(next_normal(st), m.param[1]));
        return fabs(xsub(xmul(xadd(xmul(th0, th0), th(1)), noisy), m.target[0]));
    }
    return fabs(xsub(th0, m.target[0]));
}

// ------------------------------------------------------------------ socks of Karl Broman, ref test/runtests.jl:34-44
// Spec of the draw without replacement: forward Fisher-Yates over the sorted sock list, pick j swaps
// position j with j + index(word_j, n - j); only the first m positions and the <= m displaced tail entries exist.
#define KABC_SOCKS_MAX_PICKED 32
static __device__ __noinline__ double cost_socks(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                                 uint32_t epoch, double nf, double prop) {
    const int n_picked = (int)m.param[0];
    if (!(nf >= 0.0) || !(nf <= 1e9) || nf != floor(nf) || !(prop >= 0.0 && prop <= 1.0)) return dinf();
    const long long n = (long long)nf;
    const long long n_pairs = (long long)rint(xmul(prop, floor(xdiv(nf, 2.0))));
    const int mp = n < n_picked ? (int)n : n_picked;
    long long head[KABC_SOCKS_MAX_PICKED], tpos[KABC_SOCKS_MAX_PICKED], tval[KABC_SOCKS_MAX_PICKED];
    int nt = 0;
    Stream st(rk, tag, id, epoch);
    for (int j = 0; j < mp; ++j) head[j] = j;
    for (int j = 0; j < mp; ++j) {
        long long r = j + (long long)index_of(st.next(), (uint32_t)(n - j));
        if (r < mp) { long long t = head[j]; head[j] = head[r]; head[r] = t; }
        else {
            int q = 0;
            while (q < nt && tpos[q] != r) ++q;
            if (q == nt) { tpos[nt] = r; tval[nt] = r; ++nt; }
            long long t = head[j]; head[j] = tval[q]; tval[q] = t;
        }
    }
    int lu = 0;
    for (int j = 0; j < mp; ++j) { // labels overwrite tpos (no longer needed)
        long long sidx = head[j];
        long long l = sidx < 2 * n_pairs ? sidx / 2 : sidx - n_pairs;
        bool seen = false;
        for (int q = 0; q < lu; ++q) seen |= (tpos[q] == l);
        if (!seen) tpos[lu++] = l;
    }
    const double pairs = (double)(mp - lu), odds = (double)(lu - (mp - lu));
    return xadd(fabs(xsub(pairs, m.target[0])), fabs(xsub(odds, m.target[1])));
}

// thread-per-particle dispatch.  KIND and PREC are compile-time so each kernel holds one simulator.
__device__ __forceinline__ double pushk(const DModel &m, int k, double x) { return (m.push_mask >> k) & 1u ? rint(x) : x; }

template <int KIND, int PREC, typename F>
__device__ __forceinline__ double cost_thread(const DModel &m, const RoundKeys &rk, uint32_t tag, uint32_t id,
                                              uint32_t epoch, F raw, long long &events) {
    events = 0;
    auto th = [&](int k) { return pushk(m, k, raw(k)); }; // cost(push_p(prior, theta)), ref src/smc.jl:123,176
    if (KIND == KABC_MODEL_NORMAL_MEANSTD)
        return PREC == KABC_F64 ? cost_normal_f64(m, rk, tag, id, epoch, th(0), th(1))
                                : cost_normal_f32(m, rk, tag, id, epoch, th(0), th(1));
    if (KIND == KABC_MODEL_MA2_AUTOCOV)
        return PREC == KABC_F64 ? cost_ma2_f64(m, rk, tag, id, epoch, th(0), th(1))
                                : cost_ma2_f32(m, rk, tag, id, epoch, th(0), th(1));
    if (KIND == KABC_MODEL_LV_SSA)
        return cost_lv<PREC != KABC_F64>(m, rk, tag, id, epoch, th(0), th(1), th(2), events);
    if (KIND == KABC_MODEL_DETERMINISTIC) return cost_det(m, rk, tag, id, epoch, th);
    if (KIND == KABC_MODEL_SOCKS) return cost_socks(m, rk, tag, id, epoch, th(0), th(1));
    return dnan();
}

} // namespace kabc
