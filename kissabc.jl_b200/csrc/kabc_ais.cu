// kabc_ais.cu -- sample(ApproxKernelizedPosterior(prior,cost,scale), AIS(N), Ns; ...) on device.
// Restates, per transition, KissABC.jl 3.0.1:
//   propose (mixture 4/7,2/7,1/7)   src/transition.jl:61-65
//   stretch move, a = 3             src/transition.jl:46-59
//   DE move                         src/transition.jl:2-22
//   walk move                       src/transition.jl:24-43
//   kernelized loglike              src/types.jl:51-58
//   accept (-randexp <= lW)         src/types.jl:62-75
//   init + retry budget             src/KissABC.jl:50-61
//   step / walker rotation / record src/KissABC.jl:66-80, :82-94
// Schedule: the reference moves one walker at a time (Gauss-Seidel).  The device moves a whole colour at
// once: walkers [0,h) ("red") against [h,N) ("black") and conversely, h = N/2, partners drawn from the
// complementary colour only, which keeps every simultaneous move a valid MH step (emcee's parallel
// stretch).  A sweep = two half-steps = one transition of every walker.
// Multi GPU (one process per GPU): the ensemble (theta, log-prior, second slot) is REPLICATED in every rank's peer arena;
// rank r moves N/2G red and N/2G black walkers.  Within a half-step the moving colour is read by nobody else (partners
// come from the frozen complementary colour, src/transition.jl:51-59), so an accepted walker is stored straight into
// every replica (NVLink peer stores from the accepting thread) -- this is the "all-gather of the moved colour", fused
// into the accept.  The last block of the half-step's simulate kernel meets the other ranks in a flag barrier
// (kabc_peer.cuh) and folds the counters.  Philox counters are keyed by the walker id: results do not depend on G.
#include "kabc_host.hpp"
#include "kabc_gk.cuh"
#include "kabc_peer.cuh"

namespace kabc {

enum { AIS_ERR_START_INVALID = 101, AIS_ERR_CORRECTION = 102 }; // the two error(...) of accept, src/types.jl:69-70

struct AisCtrl {
    unsigned long long accepted, cost_evals, retries;
    unsigned long long hs_accepted, hs_retries, hs_evals; // this rank, this step
    long long sweeps;
    unsigned int work_count, epoch, tk;
    int err;
};

struct AisParams {
    long long N;
    int d;
    double scale;        // kernel scale (posterior 0) or max_cost (posterior 1)
    int posterior;       // 0 ApproxKernelizedPosterior, 1 ApproxPosterior
    long long retry_cap; // per-walker attempt cap = budget + 1
    int rank, world;
};

struct AisTrace {
    unsigned char *move, *dec;
    long long *a, *b, *c;
    double *corr, *lpp, *llp, *e;
};

struct AisSlot { unsigned long long v[8]; };

struct AisBufs {
    double *th, *lp, *ll; // the local replica of the ensemble (inside the peer-visible block)
    // peer-visible block of every rank: AisSlot[2][G] | th[d][N] | lp[N] | ll[N]
    unsigned char *xb[KABC_MAX_PEERS];
    long long o_th, o_lp, o_ll;
    double *thp, *lpp, *corr;
    unsigned int *work;
    AisCtrl *ctrl;
    AisTrace tr;
    int trace_on;
};

// walkers rank r moves: a contiguous slice of each colour (colour 0 = [0,h), colour 1 = [h,N), h = N/2)
struct AisOwn { long long lo[2], n[2]; };
__host__ __device__ inline AisOwn ais_own(long long N, int rank, int world) {
    const long long h = N / 2, nb = N - h;
    AisOwn o;
    o.lo[0] = h * rank / world; o.n[0] = h * (rank + 1) / world - o.lo[0];
    o.lo[1] = h + nb * rank / world; o.n[1] = h + nb * (rank + 1) / world - o.lo[1];
    return o;
}
__device__ __forceinline__ AisSlot *aslot(const AisBufs &B, const AisParams &P, int r, int set, int src) {
    return reinterpret_cast<AisSlot *>(B.xb[r]) + set * P.world + src;
}
// a walker's new row into every other replica
__device__ __forceinline__ void ais_push_row(const AisBufs &B, const AisParams &P, long long i, double lp, double ll) {
    for (int r = 0; r < P.world; ++r) {
        if (r == P.rank) continue;
        double *pth = reinterpret_cast<double *>(B.xb[r] + B.o_th);
        for (int k = 0; k < P.d; ++k) pth[(long long)k * P.N + i] = B.th[(long long)k * P.N + i];
        reinterpret_cast<double *>(B.xb[r] + B.o_lp)[i] = lp;
        reinterpret_cast<double *>(B.xb[r] + B.o_ll)[i] = ll;
    }
}
// closes a step in the last block: cross-rank barrier + fold of the counters.  half: 0/1 = colour just moved, -1 = init
__device__ void ais_step_finish(AisBufs &B, const AisParams &P, const XPeer &x, int half) {
    AisCtrl *c = B.ctrl;
    const int set = (int)((*x.seq + 1ull) & 1ull);
    if (threadIdx.x < P.world) {
        AisSlot *s = aslot(B, P, threadIdx.x, set, P.rank);
        s->v[0] = c->hs_accepted; s->v[1] = half < 0 ? c->hs_evals : (unsigned long long)c->work_count; s->v[2] = c->hs_retries;
        s->v[3] = (unsigned long long)c->err;
    }
    if (!xbarrier(x) && threadIdx.x == 0 && !c->err) c->err = KABC_ERR_PEER;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0, work = 0, retr = 0, err = 0;
        for (int r = 0; r < P.world; ++r) {
            const AisSlot *s = aslot(B, P, P.rank, set, r);
            acc += s->v[0]; work += s->v[1]; retr += s->v[2];
            if (!err) err = s->v[3];
        }
        c->accepted += acc; c->cost_evals += work; c->retries += retr;
        if (!c->err && err) c->err = (int)err; // an error anywhere stops every rank
        c->hs_accepted = 0; c->hs_evals = 0; c->hs_retries = 0; c->work_count = 0; c->tk = 0;
        if (half >= 0) c->epoch += 1;
        if (half == 1) c->sweeps += 1;
    }
}

// ref src/types.jl:51-58 given the prior value and the cost
__device__ __forceinline__ double kernel_ll(double cost, double scale) {
    double q = xdiv(cost, scale);
    return xmul(-0.5, xmul(q, q));
}
// second slot of the walker's log-density: the kernelized log-likelihood, or -- ApproxPosterior, ref
// src/types.jl:84-91 -- the cost itself
__device__ __forceinline__ double second_slot(const AisParams &P, double cost) {
    return P.posterior == 1 ? cost : kernel_ll(cost, P.scale);
}
// ref src/types.jl:60 and :93-94
__device__ __forceinline__ bool ld_valid(const AisParams &P, double lp, double ll) {
    return P.posterior == 1 ? (dfinite(ll) && dfinite(lp)) : dfinite(xadd(lp, ll));
}

// ------------------------------------------------------------------ init with retry, ref src/KissABC.jl:50-61
// q-th walker of this rank (q < n[0]: its red slice, else its black slice)
__device__ __forceinline__ long long ais_owned_walker(const AisOwn &o, long long q) {
    return q < o.n[0] ? o.lo[0] + q : o.lo[1] + (q - o.n[0]);
}
template <int KIND, int PREC>
__global__ void __launch_bounds__(256)
k_ais_init(AisBufs B, AisParams P, XPeer x, DPriors pri, DModel m, RoundKeys rk) {
    const AisOwn own = ais_own(P.N, P.rank, P.world);
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = P.N;
    if (q < own.n[0] + own.n[1]) {
        const long long i = ais_owned_walker(own, q);
        double lp = 0, ll = 0;
        long long t = 0, evals = 0;
        bool ok = true;
        for (;; ++t) {
            Stream st(rk, ST_PRIOR, (uint32_t)i, (uint32_t)t);
            for (int k = 0; k < P.d; ++k) {
                double xv;
                ok &= prior1_sample(pri.p[k], st, xv);
                B.th[(long long)k * N + i] = xv;
            }
            const double *th = B.th;
            lp = prior_logpdf_pushed(pri, [&](int k) { return th[(long long)k * N + i]; });
            ll = P.posterior == 1 ? -lp : lp;
            if (dfinite(lp)) {
                long long ev;
                double c = cost_thread<KIND, PREC>(m, rk, ST_COST_INIT, (uint32_t)i, (uint32_t)t,
                                                   [&](int k) { return th[(long long)k * N + i]; }, ev);
                ll = second_slot(P, c);
                evals += 1;
            }
            if (ld_valid(P, lp, ll) || t >= P.retry_cap) break;
        }
        B.lp[i] = lp;
        B.ll[i] = ll;
        ais_push_row(B, P, i, lp, ll);
        if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
        if (t) atomicAdd(&B.ctrl->hs_retries, (unsigned long long)t);
        atomicAdd(&B.ctrl->hs_evals, (unsigned long long)evals);
    }
    if (last_block(&B.ctrl->tk, P.world > 1)) ais_step_finish(B, P, x, -1);
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS, (PREC == KABC_F64 ? 1 : 3))
k_ais_init_gk(AisBufs B, AisParams P, XPeer x, DPriors pri, DModel m, RoundKeys rk) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    const AisOwn own = ais_own(P.N, P.rank, P.world);
    const long long N = P.N;
    for (long long q = blockIdx.x; q < own.n[0] + own.n[1]; q += gridDim.x) {
        const long long i = ais_owned_walker(own, q);
        double lp = 0, ll = 0, xv[4];
        long long t = 0, evals = 0;
        bool ok = true;
        for (;; ++t) { // every thread draws the same prior sample (same stream), so the loop is uniform
            Stream st(rk, ST_PRIOR, (uint32_t)i, (uint32_t)t);
            for (int k = 0; k < 4; ++k) ok &= prior1_sample(pri.p[k], st, xv[k]);
            lp = prior_logpdf_pushed(pri, [&](int k) { return xv[k]; });
            ll = P.posterior == 1 ? -lp : lp;
            if (dfinite(lp)) {
                double c = cost_gk_block<PREC>(m, rk, ST_COST_INIT, (uint32_t)i, (uint32_t)t, pushk(m, 0, xv[0]), pushk(m, 1, xv[1]),
                                            pushk(m, 2, xv[2]), pushk(m, 3, xv[3]), gk_smem);
                ll = second_slot(P, c);
                evals += 1;
            }
            if (ld_valid(P, lp, ll) || t >= P.retry_cap) break;
        }
        if (threadIdx.x == 0) {
            for (int k = 0; k < 4; ++k) B.th[(long long)k * N + i] = xv[k];
            B.lp[i] = lp;
            B.ll[i] = ll;
            ais_push_row(B, P, i, lp, ll);
            if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
            if (t) atomicAdd(&B.ctrl->hs_retries, (unsigned long long)t);
            atomicAdd(&B.ctrl->hs_evals, (unsigned long long)evals);
        }
    }
    if (last_block(&B.ctrl->tk, P.world > 1)) ais_step_finish(B, P, x, -1);
}

// ------------------------------------------------------------------ propose, ref src/transition.jl:2-65
// colour range [lo,hi) moves; partners from [clo, clo+cn)
__global__ void __launch_bounds__(256)
k_ais_propose(AisBufs B, AisParams P, DPriors pri, RoundKeys rk, int colour) {
    AisCtrl *ctl = B.ctrl;
    if (ctl->err) return;
    const long long N = P.N, h = N / 2;
    const AisOwn own = ais_own(N, P.rank, P.world);
    const long long lo = own.lo[colour], hi = lo + own.n[colour];
    const long long clo = colour == 0 ? h : 0, cn = colour == 0 ? N - h : h;
    const int d = P.d;
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool push = false;
    if (i < hi) {
        const double *th = B.th;
        const uint32_t epoch = ctl->epoch;
        Stream st(rk, ST_PROPOSE, (uint32_t)i, epoch);
        long long a = i, b = i, c = i;
        double corr = 0.0;
        const uint32_t slot = index_of(st.next(), 7u); // ref :62 rand(rng,(1,1,1,1,2,2,3))
        const int move = slot < 4 ? 1 : (slot < 6 ? 2 : 3);
        if (move == 1) { // stretch
            while (a == i) a = clo + (long long)index_of(st.next(), (uint32_t)cn);
            const double u = next_uniform(st);
            const double sa = xsqrt(3.0), ra = xsqrt(xdiv(1.0, 3.0));
            const double t = xadd(xmul(u, xsub(sa, ra)), ra);
            const double Z = xmul(t, t);
            for (int k = 0; k < d; ++k) {
                const double xa = th[(long long)k * N + a], xi = th[(long long)k * N + i];
                B.thp[(long long)k * N + i] = xadd(xa, xmul(xsub(xi, xa), Z));
            }
            corr = xmul((double)(d - 1), xlog(Z));
            b = -1; c = -1;
        } else if (move == 2) { // differential evolution
            const double z0 = next_normal(st);
            const double gam = xmul(xdiv(2.38, xsqrt((double)(2 * d))), xexp(xmul(z0, 0.1)));
            while (a == i) a = clo + (long long)index_of(st.next(), (uint32_t)cn);
            while (b == a || b == i) b = clo + (long long)index_of(st.next(), (uint32_t)cn);
            for (int k = 0; k < d; ++k) {
                const double xa = th[(long long)k * N + a], xb = th[(long long)k * N + b], xi = th[(long long)k * N + i];
                const double W = xmul(xsub(xa, xb), gam);
                const double S = xadd(xadd(fabs(xsub(xa, xb)), fabs(xsub(xi, xb))), fabs(xsub(xa, xi)));
                const double T = xmul(xdiv(xmul(gam, S), 300.0), next_normal(st));
                B.thp[(long long)k * N + i] = xadd(xadd(xi, W), T);
            }
            c = -1;
        } else { // walk
            while (a == i) a = clo + (long long)index_of(st.next(), (uint32_t)cn);
            while (b == a || b == i) b = clo + (long long)index_of(st.next(), (uint32_t)cn);
            while (c == b || c == a || c == i) c = clo + (long long)index_of(st.next(), (uint32_t)cn);
            const double z1 = next_normal(st), z2 = next_normal(st), z3 = next_normal(st);
            for (int k = 0; k < d; ++k) {
                const double xa = th[(long long)k * N + a], xb = th[(long long)k * N + b], xc = th[(long long)k * N + c];
                const double xs = xdiv(xadd(xa, xadd(xb, xc)), 3.0);
                const double W = xadd(xadd(xmul(z1, xsub(xa, xs)), xmul(z2, xsub(xb, xs))), xmul(z3, xsub(xc, xs)));
                B.thp[(long long)k * N + i] = xadd(th[(long long)k * N + i], W);
            }
        }
        const double *thp = B.thp;
        const double lpp = prior_logpdf_pushed(pri, [&](int k) { return thp[(long long)k * N + i]; });
        B.lpp[i] = lpp;
        B.corr[i] = corr;
        push = dfinite(lpp);
        // ref src/types.jl:69-70: accept() raises on a non-finite correction or an invalid CURRENT state, before it looks at
        // the proposal (so also for proposals outside the prior support)
        if (!dfinite(corr)) ctl->err = AIS_ERR_CORRECTION;
        else if (!ld_valid(P, B.lp[i], B.ll[i])) ctl->err = AIS_ERR_START_INVALID;
        if (B.trace_on) {
            B.tr.move[i] = (unsigned char)move; B.tr.a[i] = a; B.tr.b[i] = b; B.tr.c[i] = c; B.tr.corr[i] = corr;
            B.tr.lpp[i] = lpp; B.tr.llp[i] = P.posterior == 1 ? -lpp : lpp; B.tr.e[i] = dnan(); B.tr.dec[i] = 0;
        }
    }
    // block-aggregated append to the work list: one global atomic per CTA
    __shared__ unsigned int s_cnt[8], s_base;
    const unsigned int ball = __ballot_sync(0xffffffffu, push);
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = __popc(ball);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const unsigned int v = s_cnt[w]; s_cnt[w] = tot; tot += v; }
        s_base = tot ? atomicAdd(&ctl->work_count, tot) : 0u;
    }
    __syncthreads();
    if (push) B.work[s_base + s_cnt[warp] + __popc(ball & ((1u << lane) - 1u))] = (unsigned int)i;
}

// ref src/types.jl:62-75 + src/transition.jl:76-79
__device__ __forceinline__ unsigned int ais_accept(AisBufs &B, const AisParams &P, const RoundKeys &rk, long long i,
                                                   uint32_t epoch, double cost) {
    const long long N = P.N;
    const double lpp = B.lpp[i];
    const double llp = second_slot(P, cost);
    int dec = 0;
    double e = dnan();
    if (ld_valid(P, lpp, llp)) {
        Stream sa(rk, ST_ACCEPT, (uint32_t)i, epoch);
        e = next_exp(sa);
        if (P.posterior == 1) { // ref src/types.jl:101-103: (-randexp <= lW) && lW2 >= 0
            const double lW = xsub(xadd(B.corr[i], lpp), B.lp[i]);
            const double lW2 = xsub(fmax(P.scale, B.ll[i]), llp);
            dec = ((-e <= lW) && lW2 >= 0.0) ? 2 : 1;
        } else {                // ref src/types.jl:73-74
            const double lW = xsub(xadd(B.corr[i], xadd(lpp, llp)), xadd(B.lp[i], B.ll[i]));
            dec = (-e <= lW) ? 2 : 1;
        }
    }
    if (dec == 2) {
        for (int k = 0; k < P.d; ++k) B.th[(long long)k * N + i] = B.thp[(long long)k * N + i];
        B.lp[i] = lpp;
        B.ll[i] = llp;
        ais_push_row(B, P, i, lpp, llp); // the moved walker into every replica (nobody reads its colour in this half-step)
    }
    if (B.trace_on) { B.tr.llp[i] = llp; B.tr.e[i] = e; B.tr.dec[i] = (unsigned char)dec; }
    return dec == 2;
}

template <int KIND, int PREC>
__global__ void __launch_bounds__(256) k_ais_simulate(AisBufs B, AisParams P, XPeer x, DModel m, RoundKeys rk, int colour) {
    AisCtrl *ctl = B.ctrl;
    if (ctl->err == KABC_ERR_PEER) return;
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nwork = ctl->err ? 0u : ctl->work_count; // after an error: no work, but still meet the other ranks
    if ((w & ~31u) < nwork) {
        unsigned int acc = 0;
        if (w < nwork) {
            const long long i = B.work[w];
            const long long N = P.N;
            const double *thp = B.thp;
            long long ev;
            const uint32_t epoch = ctl->epoch;
            double c = cost_thread<KIND, PREC>(m, rk, ST_COST, (uint32_t)i, epoch, [&](int k) { return thp[(long long)k * N + i]; }, ev);
            acc = ais_accept(B, P, rk, i, epoch, c);
        }
        unsigned int nacc = __popc(__ballot_sync(0xffffffffu, acc));
        if ((threadIdx.x & 31) == 0 && nacc) atomicAdd(&ctl->hs_accepted, (unsigned long long)nacc);
    }
    if (last_block(&ctl->tk, P.world > 1)) ais_step_finish(B, P, x, colour);
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS, (PREC == KABC_F64 ? 1 : 3)) k_ais_simulate_gk(AisBufs B, AisParams P, XPeer x, DModel m, RoundKeys rk, int colour) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    AisCtrl *ctl = B.ctrl;
    if (ctl->err == KABC_ERR_PEER) return;
    const unsigned int nwork = ctl->err ? 0u : ctl->work_count;
    const long long N = P.N;
    const uint32_t epoch = ctl->epoch;
    for (unsigned int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const long long i = B.work[w];
        const double *thp = B.thp;
        double c = cost_gk_block<PREC>(m, rk, ST_COST, (uint32_t)i, epoch, pushk(m, 0, thp[i]), pushk(m, 1, thp[N + i]),
                                    pushk(m, 2, thp[2 * N + i]), pushk(m, 3, thp[3 * N + i]), gk_smem);
        if (threadIdx.x == 0 && ais_accept(B, P, rk, i, epoch, c)) atomicAdd(&ctl->hs_accepted, 1ull);
    }
    if (last_block(&ctl->tk, P.world > 1)) ais_step_finish(B, P, x, colour);
}

__global__ void k_ais_reset(AisBufs B) {
    AisCtrl *c = B.ctrl;
    memset(c, 0, sizeof(AisCtrl));
}

// bundle_samples, ref src/KissABC.jl:78,90-93: saved sample m is walker w_m of the current ensemble
__global__ void k_ais_record(AisBufs B, AisParams P, double *out, long long Ns, long long m0, long long m1,
                             long long discard, long long thinning, uint32_t push_mask) {
    long long m = m0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= m1) return;
    const long long target = discard + m * thinning;
    const long long w = target > 0 ? (target - 1) % P.N : P.N - 1;
    for (int k = 0; k < P.d; ++k) { // ref src/KissABC.jl:78: the recorded sample is push_p(model, sample[i])
        double v = B.th[(long long)k * P.N + w];
        out[(long long)k * Ns + m] = (push_mask >> k) & 1u ? rint(v) : v;
    }
}

} // namespace kabc

using namespace kabc;

struct kabc_ais {
    kabc_ctx *ctx = nullptr;
    DPriors pri;
    DModel model;
    AisParams P;
    kabc_ais_config_t cfg;
    AisBufs B;
    XPeer X;
    DevBuf<unsigned char> xlocal; // the peer-visible block when there are no peers
    bool in_arena = false;
    DevBuf<double> thp, lpp, corr;
    DevBuf<unsigned int> work;
    DevBuf<AisCtrl> ctrl;
    DevBuf<unsigned char> tmove, tdec;
    DevBuf<long long> ta, tb, tc;
    DevBuf<double> tcorr, tlpp, tllp, te;
    AisCtrl *h_ctrl = nullptr;
    bool inited = false;
    long long launches = 0;
    cudaGraphExec_t sweep_graph = nullptr; // one red/black sweep (4 kernels) captured once
    bool graph_ok = true;
};

#define AIS_LAUNCHED(s) do { (s)->launches += 1; (s)->ctx->launches += 1; } while (0)

static inline size_t ais_align256(size_t b) { return (b + 255) & ~(size_t)255; }
struct AisLayout { size_t o_th, o_lp, o_ll, bytes; };
static AisLayout ais_layout(long long N, int d, int world) {
    AisLayout L;
    size_t o = ais_align256(sizeof(AisSlot) * 2 * (size_t)world);
    L.o_th = o; o += ais_align256((size_t)N * d * 8);
    L.o_lp = o; o += ais_align256((size_t)N * 8);
    L.o_ll = o; o += ais_align256((size_t)N * 8);
    L.bytes = o;
    return L;
}

static int ais_read_ctrl(kabc_ais *s) {
    KABC_CUDA_TRY(cudaMemcpyAsync(s->h_ctrl, s->B.ctrl, sizeof(AisCtrl), cudaMemcpyDeviceToHost, s->ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    return KABC_OK;
}
static int ais_ctrl_error(kabc_ais *s) {
    switch (s->h_ctrl->err) {
    case 0: return KABC_OK;
    case AIS_ERR_START_INVALID: return set_error(KABC_ERR_STATE, "starting sample invalid.");  // ref src/types.jl:70
    case AIS_ERR_CORRECTION: return set_error(KABC_ERR_STATE, "ld_correction is invalid");     // ref src/types.jl:69
    case KABC_ERR_PEER: return set_error(KABC_ERR_PEER, "a rank of the job did not reach a cross-rank barrier in time");
    default: return set_error(KABC_ERR_INVALID_ARG, "prior sampling failed (truncation too extreme)");
    }
}

static int ais_gk_grid(kabc_ais *s, long long n, size_t &smem) {
    smem = gk_smem_bytes(s->model.n_draws, s->model.precision);
    long long cap = (long long)s->ctx->sm_count * gk_blocks_per_sm(s->model.n_draws, s->model.precision);
    if (n < 1) n = 1;
    return (int)(n < cap ? n : cap);
}

template <int KIND>
static void ais_launch_init_t(kabc_ais *s, long long n_own) {
    const unsigned blocks = (unsigned)((n_own + 255) / 256 > 0 ? (n_own + 255) / 256 : 1);
    if (s->model.precision == KABC_F64)
        k_ais_init<KIND, KABC_F64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->X, s->pri, s->model, s->ctx->rk);
    else
        k_ais_init<KIND, KABC_F32_ACC64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->X, s->pri, s->model, s->ctx->rk);
    AIS_LAUNCHED(s);
}

template <int KIND>
static void ais_launch_sim_t(kabc_ais *s, long long n, int colour) {
    const unsigned blocks = (unsigned)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1);
    if (s->model.precision == KABC_F64)
        k_ais_simulate<KIND, KABC_F64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->X, s->model, s->ctx->rk, colour);
    else
        k_ais_simulate<KIND, KABC_F32_ACC64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->X, s->model, s->ctx->rk, colour);
    AIS_LAUNCHED(s);
}

static int ais_enqueue_half(kabc_ais *s, int colour) {
    NvtxRange nv(colour == 0 ? "kabc:ais:half-step red" : "kabc:ais:half-step black");
    kabc_ctx *ctx = s->ctx;
    const AisOwn own = ais_own(s->P.N, ctx->rank, ctx->world);
    const long long n = own.n[colour];
    k_ais_propose<<<(unsigned)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1), 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk, colour);
    AIS_LAUNCHED(s);
    switch (s->model.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: ais_launch_sim_t<KABC_MODEL_NORMAL_MEANSTD>(s, n, colour); break;
    case KABC_MODEL_MA2_AUTOCOV: ais_launch_sim_t<KABC_MODEL_MA2_AUTOCOV>(s, n, colour); break;
    case KABC_MODEL_LV_SSA: ais_launch_sim_t<KABC_MODEL_LV_SSA>(s, n, colour); break;
    case KABC_MODEL_DETERMINISTIC: ais_launch_sim_t<KABC_MODEL_DETERMINISTIC>(s, n, colour); break;
    case KABC_MODEL_SOCKS: ais_launch_sim_t<KABC_MODEL_SOCKS>(s, n, colour); break;
    case KABC_MODEL_GK_OCTILE: {
        size_t smem;
        int grid = ais_gk_grid(s, n, smem);
        if (s->model.precision == KABC_F64)
            k_ais_simulate_gk<KABC_F64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->X, s->model, ctx->rk, colour);
        else
            k_ais_simulate_gk<KABC_F32_ACC64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->X, s->model, ctx->rk, colour);
        AIS_LAUNCHED(s);
        break;
    }
    }
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

// one red/black sweep = 2 half-steps = 4 kernels, replayed from a CUDA graph (small ensembles are launch bound)
static int ais_launch_sweep(kabc_ais *s) {
    NvtxRange nv("kabc:ais:sweep");
    kabc_ctx *ctx = s->ctx;
    static const bool env_off = [] { const char *e = getenv("KABC_NO_GRAPH"); return e && e[0] == '1'; }();
    if (!s->graph_ok || env_off) {
        if (int rc = ais_enqueue_half(s, 0)) return rc;
        return ais_enqueue_half(s, 1);
    }
    if (!s->sweep_graph) {
        const long long l0 = s->launches, c0 = ctx->launches;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        int rc = KABC_OK;
        if (e == cudaSuccess) {
            rc = ais_enqueue_half(s, 0);
            if (!rc) rc = ais_enqueue_half(s, 1);
            e = cudaStreamEndCapture(ctx->stream, &g);
        }
        s->launches = l0; ctx->launches = c0;
        if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&s->sweep_graph, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess || rc) {
            cudaGetLastError();
            s->sweep_graph = nullptr;
            s->graph_ok = false;
            if (int rc2 = ais_enqueue_half(s, 0)) return rc2;
            return ais_enqueue_half(s, 1);
        }
    }
    KABC_CUDA_TRY(cudaGraphLaunch(s->sweep_graph, ctx->stream));
    s->launches += 4; ctx->launches += 4;
    return KABC_OK;
}

extern "C" {

uint64_t kabc_ais_arena_bytes(int64_t nwalkers, int d, int world) {
    if (world < 1 || nwalkers < 1 || d < 1) return 0;
    return (uint64_t)ais_layout(nwalkers, d, world).bytes + 4096;
}

int kabc_ais_create(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                    const kabc_ais_config_t *cfg, kabc_ais_t **out) {
    if (!ctx || !cfg || !out) return set_error(KABC_ERR_INVALID_ARG, "NULL argument");
    DPriors pri;
    DModel m;
    if (int rc = ingest_priors(prior, d, pri)) return rc;
    // ref src/KissABC.jl:43-48
    if (cfg->nwalkers < d + 5)
        return set_error(KABC_ERR_INVALID_ARG, "nparticles = %lld is insufficient, set number of particles in AIS(.) atleast to %d",
                         (long long)cfg->nwalkers, d + 5);
    if (cfg->nwalkers > 0x7FFFFFFFll) return set_error(KABC_ERR_INVALID_ARG, "nwalkers must be < 2^31");
    if (cfg->posterior != 0 && cfg->posterior != 1) return set_error(KABC_ERR_INVALID_ARG, "posterior must be 0 (kernelized) or 1 (hard threshold)");
    if (cfg->posterior == 0 && !(cfg->scale > 0)) return set_error(KABC_ERR_INVALID_ARG, "kernel scale (target_average_cost) must be > 0");
    if (cfg->nsamples < 0 || cfg->ntransitions < 1 || cfg->discard_initial < 0 || cfg->thinning < 1 || cfg->retry_sampling < 0)
        return set_error(KABC_ERR_INVALID_ARG, "bad AIS configuration");
    if (int rc = ingest_model(model, d, m)) return rc;
    m.push_mask = push_mask_of(pri);
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    kabc_ais *s = new kabc_ais();
    s->ctx = ctx; s->pri = pri; s->model = m; s->cfg = *cfg;
    const long long N = cfg->nwalkers;
    s->P.N = N; s->P.d = d; s->P.scale = cfg->scale; s->P.posterior = cfg->posterior; s->P.retry_cap = cfg->retry_sampling * N + 1;
    s->P.rank = ctx->rank; s->P.world = ctx->world;
    const size_t nd = (size_t)N * d;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    const AisLayout L = ais_layout(N, d, ctx->world);
    memset(s->B.xb, 0, sizeof s->B.xb);
    if (ctx->world > 1) {
        size_t off = 0;
        if (int rc = arena_alloc(ctx, L.bytes, &off)) { delete s; return rc; }
        s->in_arena = true;
        for (int r = 0; r < ctx->world; ++r) s->B.xb[r] = (unsigned char *)ctx->arena_map[r] + off;
    } else {
        A(s->xlocal.alloc(ctx, L.bytes));
        s->B.xb[0] = s->xlocal.p;
    }
    s->X = make_xpeer(ctx);
    s->B.o_th = (long long)L.o_th; s->B.o_lp = (long long)L.o_lp; s->B.o_ll = (long long)L.o_ll;
    A(s->thp.alloc(ctx, nd)); A(s->lpp.alloc(ctx, N));
    A(s->corr.alloc(ctx, N)); A(s->work.alloc(ctx, N)); A(s->ctrl.alloc(ctx, 1));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&s->h_ctrl, sizeof(AisCtrl));
    if (e != cudaSuccess) {
        if (s->in_arena) arena_release(ctx);
        delete s;
        return set_error(KABC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    memset(s->h_ctrl, 0, sizeof(AisCtrl));
    unsigned char *own = s->B.xb[ctx->rank];
    s->B.th = reinterpret_cast<double *>(own + L.o_th); s->B.lp = reinterpret_cast<double *>(own + L.o_lp);
    s->B.ll = reinterpret_cast<double *>(own + L.o_ll);
    s->B.thp = s->thp.p; s->B.lpp = s->lpp.p; s->B.corr = s->corr.p; s->B.work = s->work.p; s->B.ctrl = s->ctrl.p;
    memset(&s->B.tr, 0, sizeof s->B.tr);
    s->B.trace_on = 0;
    cudaError_t e2 = cudaMemsetAsync(s->ctrl.p, 0, sizeof(AisCtrl), ctx->stream);
    if (e2 == cudaSuccess && m.kind == KABC_MODEL_GK_OCTILE) {
        const int smem = (int)gk_smem_bytes(m.n_draws, m.precision);
        if (m.precision == KABC_F64) {
            e2 = cudaFuncSetAttribute(k_ais_init_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_ais_simulate_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        } else {
            e2 = cudaFuncSetAttribute(k_ais_init_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_ais_simulate_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        }
    }
    if (e2 != cudaSuccess) {
        kabc_ais_destroy(s);
        return set_error(KABC_ERR_CUDA, "ais handle setup failed: %s", cudaGetErrorString(e2));
    }
    *out = s;
    return KABC_OK;
}

int kabc_ais_destroy(kabc_ais_t *s) {
    if (!s) return KABC_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->sweep_graph) cudaGraphExecDestroy(s->sweep_graph);
    if (s->in_arena) arena_release(s->ctx);
    if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
    delete s;
    return KABC_OK;
}

int kabc_ais_init(kabc_ais_t *s) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "ais is NULL");
    kabc_ctx *ctx = s->ctx;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    k_ais_reset<<<1, 1, 0, ctx->stream>>>(s->B);
    AIS_LAUNCHED(s);
    const AisOwn own = ais_own(s->P.N, ctx->rank, ctx->world);
    const long long n_own = own.n[0] + own.n[1];
    switch (s->model.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: ais_launch_init_t<KABC_MODEL_NORMAL_MEANSTD>(s, n_own); break;
    case KABC_MODEL_MA2_AUTOCOV: ais_launch_init_t<KABC_MODEL_MA2_AUTOCOV>(s, n_own); break;
    case KABC_MODEL_LV_SSA: ais_launch_init_t<KABC_MODEL_LV_SSA>(s, n_own); break;
    case KABC_MODEL_DETERMINISTIC: ais_launch_init_t<KABC_MODEL_DETERMINISTIC>(s, n_own); break;
    case KABC_MODEL_SOCKS: ais_launch_init_t<KABC_MODEL_SOCKS>(s, n_own); break;
    case KABC_MODEL_GK_OCTILE: {
        size_t smem;
        int grid = ais_gk_grid(s, n_own, smem);
        if (s->model.precision == KABC_F64)
            k_ais_init_gk<KABC_F64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->X, s->pri, s->model, ctx->rk);
        else
            k_ais_init_gk<KABC_F32_ACC64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->X, s->pri, s->model, ctx->rk);
        AIS_LAUNCHED(s);
        break;
    }
    }
    KABC_CUDA_TRY(cudaGetLastError());
    if (int rc = ais_read_ctrl(s)) return rc;
    if (int rc = ais_ctrl_error(s)) return rc;
    if ((long long)s->h_ctrl->retries > s->cfg.retry_sampling * s->P.N)
        return set_error(KABC_ERR_RETRY_BUDGET, "Prior leads to \xe2\x88\x9e costs too often, tune the prior or increase `retry_sampling`.");
    s->inited = true;
    return KABC_OK;
}

int kabc_ais_sweep(kabc_ais_t *s, int nsweeps, float *out_ms) {
    if (!s || nsweeps < 0) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    if (!s->inited) return set_error(KABC_ERR_STATE, "kabc_ais_init must be called first");
    if (s->P.N / 2 < 3) return set_error(KABC_ERR_INVALID_ARG, "red/black AIS needs >= 3 walkers per colour");
    kabc_ctx *ctx = s->ctx;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    KABC_CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int r = 0; r < nsweeps; ++r)
        if (int rc = ais_launch_sweep(s)) return rc;
    KABC_CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    if (int rc = ais_read_ctrl(s)) return rc;
    if (out_ms) KABC_CUDA_TRY(cudaEventElapsedTime(out_ms, ctx->ev0, ctx->ev1));
    return ais_ctrl_error(s);
}

int kabc_ais_get_state(kabc_ais_t *s, double *theta, double *lp, double *ll) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "ais is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const size_t N = (size_t)s->P.N;
    cudaStream_t st = s->ctx->stream;
    if (theta) KABC_CUDA_TRY(cudaMemcpyAsync(theta, s->B.th, 8 * N * s->P.d, cudaMemcpyDeviceToHost, st));
    if (lp) KABC_CUDA_TRY(cudaMemcpyAsync(lp, s->B.lp, 8 * N, cudaMemcpyDeviceToHost, st));
    if (ll) KABC_CUDA_TRY(cudaMemcpyAsync(ll, s->B.ll, 8 * N, cudaMemcpyDeviceToHost, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    return KABC_OK;
}

// every rank passes the whole ensemble (the replicas stay identical)
int kabc_ais_set_state(kabc_ais_t *s, const double *theta, const double *lp, const double *ll) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "ais is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const size_t N = (size_t)s->P.N;
    cudaStream_t st = s->ctx->stream;
    if (theta) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.th, theta, 8 * N * s->P.d, cudaMemcpyHostToDevice, st));
    if (lp) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.lp, lp, 8 * N, cudaMemcpyHostToDevice, st));
    if (ll) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.ll, ll, 8 * N, cudaMemcpyHostToDevice, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    s->inited = true;
    return KABC_OK;
}

int kabc_ais_get_counters(kabc_ais_t *s, int64_t *cost_evals, int64_t *accepted, int64_t *sweeps, int64_t *retries) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "ais is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (int rc = ais_read_ctrl(s)) return rc;
    if (cost_evals) *cost_evals = (int64_t)s->h_ctrl->cost_evals;
    if (accepted) *accepted = (int64_t)s->h_ctrl->accepted;
    if (sweeps) *sweeps = s->h_ctrl->sweeps;
    if (retries) *retries = (int64_t)s->h_ctrl->retries;
    return KABC_OK;
}

int64_t kabc_ais_kernel_launches(kabc_ais_t *s) { return s ? s->launches : -1; }

int kabc_ais_trace_enable(kabc_ais_t *s, int on) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "ais is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    KABC_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    if (on && !s->ta.p) {
        const size_t N = (size_t)s->P.N;
        KABC_CUDA_TRY(s->tmove.alloc(N)); KABC_CUDA_TRY(s->tdec.alloc(N)); KABC_CUDA_TRY(s->ta.alloc(N));
        KABC_CUDA_TRY(s->tb.alloc(N)); KABC_CUDA_TRY(s->tc.alloc(N)); KABC_CUDA_TRY(s->tcorr.alloc(N));
        KABC_CUDA_TRY(s->tlpp.alloc(N)); KABC_CUDA_TRY(s->tllp.alloc(N)); KABC_CUDA_TRY(s->te.alloc(N));
        s->B.tr.move = s->tmove.p; s->B.tr.dec = s->tdec.p; s->B.tr.a = s->ta.p; s->B.tr.b = s->tb.p; s->B.tr.c = s->tc.p;
        s->B.tr.corr = s->tcorr.p; s->B.tr.lpp = s->tlpp.p; s->B.tr.llp = s->tllp.p; s->B.tr.e = s->te.p;
    }
    s->B.trace_on = on ? 1 : 0;
    if (s->sweep_graph) { cudaGraphExecDestroy(s->sweep_graph); s->sweep_graph = nullptr; } // the captured launches hold AisBufs by value
    return KABC_OK;
}

int kabc_ais_get_trace(kabc_ais_t *s, uint8_t *move, int64_t *a, int64_t *b, int64_t *c, double *corr, double *theta_p,
                       double *lp_p, double *ll_p, double *e, uint8_t *decision) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "ais is NULL");
    if (!s->ta.p) return set_error(KABC_ERR_STATE, "trace was never enabled");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const size_t N = (size_t)s->P.N;
    cudaStream_t st = s->ctx->stream;
    if (move) KABC_CUDA_TRY(cudaMemcpyAsync(move, s->tmove.p, N, cudaMemcpyDeviceToHost, st));
    if (a) KABC_CUDA_TRY(cudaMemcpyAsync(a, s->ta.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (b) KABC_CUDA_TRY(cudaMemcpyAsync(b, s->tb.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (c) KABC_CUDA_TRY(cudaMemcpyAsync(c, s->tc.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (corr) KABC_CUDA_TRY(cudaMemcpyAsync(corr, s->tcorr.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (theta_p) KABC_CUDA_TRY(cudaMemcpyAsync(theta_p, s->thp.p, 8 * N * s->P.d, cudaMemcpyDeviceToHost, st));
    if (lp_p) KABC_CUDA_TRY(cudaMemcpyAsync(lp_p, s->tlpp.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (ll_p) KABC_CUDA_TRY(cudaMemcpyAsync(ll_p, s->tllp.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (e) KABC_CUDA_TRY(cudaMemcpyAsync(e, s->te.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (decision) KABC_CUDA_TRY(cudaMemcpyAsync(decision, s->tdec.p, N, cudaMemcpyDeviceToHost, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    return KABC_OK;
}

int kabc_ais_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                 const kabc_ais_config_t *cfg, double *out_samples, int64_t *out_cost_evals, int64_t *out_accepted) {
    if (!out_samples) return set_error(KABC_ERR_INVALID_ARG, "out_samples is NULL");
    kabc_ais *s = nullptr;
    if (int rc = kabc_ais_create(ctx, prior, d, model, cfg, &s)) return rc;
    int rc = kabc_ais_init(s);
    const long long N = s->P.N, Ns = cfg->nsamples;
    DevBuf<double> dout;
    if (!rc && Ns > 0) {
        cudaError_t e = dout.alloc((size_t)Ns * d);
        if (e != cudaSuccess) rc = set_error(KABC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    auto need_of = [&](long long m) {
        long long target = cfg->discard_initial + m * cfg->thinning;
        return target > 0 ? (target + N - 1) / N : 0ll;
    };
    long long rounds = 0, m = 0;
    if (!rc && Ns > 0 && need_of(Ns - 1) > 0 && N / 2 < 3) rc = set_error(KABC_ERR_INVALID_ARG, "red/black AIS needs >= 3 walkers per colour");
    while (!rc && m < Ns) {
        const long long need = need_of(m);
        while (!rc && rounds < need) { // one round = ntransitions sweeps of the whole ensemble
            for (long long r = 0; !rc && r < cfg->ntransitions; ++r) rc = ais_launch_sweep(s);
            ++rounds;
        }
        long long m1 = m;
        while (m1 < Ns && need_of(m1) == need) ++m1;
        if (!rc) {
            k_ais_record<<<(unsigned)((m1 - m + 255) / 256), 256, 0, ctx->stream>>>(s->B, s->P, dout.p, Ns, m, m1, cfg->discard_initial, cfg->thinning, s->model.push_mask);
            AIS_LAUNCHED(s);
        }
        m = m1;
    }
    if (!rc && Ns > 0) {
        cudaError_t e = cudaMemcpyAsync(out_samples, dout.p, sizeof(double) * (size_t)Ns * d, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = set_error(KABC_ERR_CUDA, "copy of samples failed: %s", cudaGetErrorString(e));
    }
    if (!rc) rc = kabc_ais_get_counters(s, out_cost_evals, out_accepted, nullptr, nullptr);
    if (!rc) rc = ais_ctrl_error(s); // ref src/types.jl:69-70
    std::string keep = g_last_error;
    kabc_ais_destroy(s);
    g_last_error = keep;
    return rc;
}

} // extern "C"
