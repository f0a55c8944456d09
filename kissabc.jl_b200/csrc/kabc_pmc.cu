// kabc_pmc.cu -- the two other population samplers of the reference, ABCDE (src/smc.jl:352-428) and pfilter
// (src/smc.jl:275-345), on the building blocks of the smc path: SoA FP64 state, Philox streams keyed by
// (tag, particle, epoch), prior kernels, registered simulators.
//
// Both share one schedule: a PENDING list of particles -> a propose kernel (partner picks, proposal, prior
// Metropolis pre-test) that appends the survivors to an EVAL list -> the list-driven simulator launch
// (eval_cost_list_device) -> an accept kernel that commits in place or re-queues.  Proposals only read rows that no
// accept of the same round writes (ABCDE: all proposals of a generation are formed before any accept, the Jacobi update
// of ref :379-381; pfilter: partners come from the particles under the quantile, ref :309-311), so the state needs no
// second copy.  "uniform choice out of a subset" (ref :393 and :309) is taken over the subset in (cost, index) order,
// which both algorithms get from one bitonic sort of (order-preserving key, index) pairs per generation / iteration.
#include "../../include/kissabc_cuda.h"
#include "kabc_host.hpp"

namespace kabc {

struct PmcCtrl {
    unsigned int n_eval, n_pend2, n_ok;
    int err, stop;
    long long gens;
    unsigned long long evals;
    double eps, eps_pop;
};

struct PmcBufs {
    double *th, *lp, *C;    // state: particles (SoA d x N), log-prior of push_p(particle), cost
    double *thp, *lpp, *Cp; // proposals
    double *thr;            // ABCDE: acceptance threshold max(eps, cost_i) of the proposal
    unsigned long long *keys;
    unsigned int *sidx;     // sorted (cost, index) order, padded to a power of two
    unsigned int *list, *pendA, *pendB;
    PmcCtrl *ctrl;
};

__device__ __forceinline__ unsigned long long pmc_key(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double pmc_unkey(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}
// first position in the sorted keys[0..n) whose key exceeds k
__device__ __forceinline__ unsigned int pmc_upper_bound(const unsigned long long *keys, unsigned int n, unsigned long long k) {
    unsigned int lo = 0, hi = n;
    while (lo < hi) {
        const unsigned int mid = (lo + hi) >> 1;
        if (keys[mid] <= k) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------ (cost, index) sort
__global__ void k_pmc_sort_fill(const double *C, unsigned int N, unsigned int M, unsigned long long *keys, unsigned int *sidx) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    keys[i] = i < N ? pmc_key(C[i]) : ~0ull;
    sidx[i] = i;
}
__device__ __forceinline__ bool pmc_pair_gt(unsigned long long ka, unsigned int ia, unsigned long long kb, unsigned int ib) {
    return ka > kb || (ka == kb && ia > ib);
}
constexpr int PMC_SORT_BLOCK = 4096; // pairs one CTA sorts in shared memory
__global__ void __launch_bounds__(1024) k_pmc_sort_block(unsigned long long *keys, unsigned int *sidx, unsigned int M) {
    __shared__ unsigned long long sk[PMC_SORT_BLOCK];
    __shared__ unsigned int si[PMC_SORT_BLOCK];
    for (unsigned int i = threadIdx.x; i < M; i += blockDim.x) { sk[i] = keys[i]; si[i] = sidx[i]; }
    __syncthreads();
    for (unsigned int k = 2; k <= M; k <<= 1)
        for (unsigned int j = k >> 1; j > 0; j >>= 1) {
            for (unsigned int i = threadIdx.x; i < M; i += blockDim.x) {
                const unsigned int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    if (pmc_pair_gt(sk[i], si[i], sk[l], si[l]) == up) {
                        const unsigned long long tk = sk[i]; sk[i] = sk[l]; sk[l] = tk;
                        const unsigned int ti = si[i]; si[i] = si[l]; si[l] = ti;
                    }
                }
            }
            __syncthreads();
        }
    for (unsigned int i = threadIdx.x; i < M; i += blockDim.x) { keys[i] = sk[i]; sidx[i] = si[i]; }
}
__global__ void k_pmc_sort_step(unsigned long long *keys, unsigned int *sidx, unsigned int M, unsigned int j, unsigned int k) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const unsigned int l = i ^ j;
    if (l <= i) return;
    const bool up = (i & k) == 0;
    const unsigned long long ka = keys[i], kb = keys[l];
    const unsigned int ia = sidx[i], ib = sidx[l];
    if (pmc_pair_gt(ka, ia, kb, ib) == up) { keys[i] = kb; keys[l] = ka; sidx[i] = ib; sidx[l] = ia; }
}

// ------------------------------------------------------------------ init, ref :281-297 and :355-371
__global__ void k_pmc_iota(unsigned int *v, unsigned int n) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}
__global__ void k_pmc_reset_round(PmcCtrl *c) { c->n_eval = 0; c->n_pend2 = 0; }
// try t of the particles still without finite cost and log-prior: stream (PRIOR, i, t)
__global__ void k_pmc_init_propose(PmcBufs B, long long N, DPriors pri, RoundKeys rk, const unsigned int *pend, unsigned int n_pend,
                                   unsigned int *pend2, uint32_t t) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_pend) return;
    const long long i = pend[w];
    Stream st(rk, ST_PRIOR, (uint32_t)i, t);
    bool ok = true;
    for (int k = 0; k < pri.d; ++k) {
        double x;
        ok &= prior1_sample(pri.p[k], st, x);
        B.th[(long long)k * N + i] = x;
    }
    if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
    const double *th = B.th;
    const double l = prior_logpdf_pushed(pri, [&](int k) { return th[(long long)k * N + i]; });
    B.lp[i] = l;
    if (dfinite(l)) B.list[atomicAdd(&B.ctrl->n_eval, 1u)] = (unsigned int)i;
    else pend2[atomicAdd(&B.ctrl->n_pend2, 1u)] = (unsigned int)i;
}
__global__ void k_pmc_init_accept(PmcBufs B, unsigned int *pend2) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n = B.ctrl->n_eval;
    if (w == 0) B.ctrl->evals += n;
    if (w >= n) return;
    const unsigned int i = B.list[w];
    const double c = B.Cp[i];
    if (dfinite(c)) B.C[i] = c;
    else pend2[atomicAdd(&B.ctrl->n_pend2, 1u)] = i;
}

// ------------------------------------------------------------------ ABCDE generation, ref :376-417
struct AbcdeParams {
    double eps_target, alpha, gamma;
    int earlystop;
};
__global__ void k_abcde_pre(PmcBufs B, unsigned int N, AbcdeParams A) {
    PmcCtrl *c = B.ctrl;
    c->n_eval = 0;
    if (c->stop) return;
    const double eps_l = pmc_unkey(B.keys[0]), eps_h = pmc_unkey(B.keys[N - 1]); // ref :382 extrema(costs)
    if (A.earlystop && eps_h <= A.eps_target) { c->stop = 1; return; }
    c->eps_pop = fmax(A.eps_target, xadd(eps_l, xmul(A.alpha, xsub(eps_h, eps_l))));
    c->gens += 1;
}
__global__ void __launch_bounds__(256) k_abcde_propose(PmcBufs B, long long N, DPriors pri, RoundKeys rk, AbcdeParams A) {
    const PmcCtrl *c = B.ctrl;
    if (c->stop) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double Di = B.C[i];
    if (A.earlystop && Di <= A.eps_target) return;
    Stream st(rk, ST_PROPOSE, (uint32_t)i, (uint32_t)c->gens);
    const double eps = Di <= A.eps_target ? A.eps_target : c->eps_pop;
    long long s = i;
    if (Di > eps) { // ref :393: a uniformly chosen particle that is at least as good
        const unsigned int ub = pmc_upper_bound(B.keys, (unsigned int)N, pmc_key(Di));
        s = B.sidx[index_of(st.next(), ub)];
    }
    long long a = s, b;
    while (a == s) a = (long long)index_of(st.next(), (uint32_t)N);
    b = a;
    while (b == a || b == s) b = (long long)index_of(st.next(), (uint32_t)N);
    for (int k = 0; k < pri.d; ++k) {
        const double *t = B.th + (long long)k * N;
        B.thp[(long long)k * N + i] = xadd(t[s], xmul(xsub(t[a], t[b]), A.gamma)); // ref :403
    }
    const double *thp = B.thp;
    const double l = prior_logpdf_pushed(pri, [&](int k) { return thp[(long long)k * N + i]; });
    const double w = xsub(l, B.lp[i]);
    if (xlog(u01(st.next())) > fmin(0.0, w)) return; // ref :406
    B.lpp[i] = l;
    B.thr[i] = fmax(eps, Di);
    B.list[atomicAdd(&B.ctrl->n_eval, 1u)] = (unsigned int)i;
}
__global__ void k_abcde_accept(PmcBufs B, long long N, int d) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n = B.ctrl->n_eval;
    if (w == 0) B.ctrl->evals += n; // ref :407 nsims
    if (w >= n) return;
    const long long i = B.list[w];
    const double dp = B.Cp[i];
    if (dp <= B.thr[i]) { // ref :409
        B.C[i] = dp;
        for (int k = 0; k < d; ++k) B.th[(long long)k * N + i] = B.thp[(long long)k * N + i];
        B.lp[i] = B.lpp[i];
    }
}

// ------------------------------------------------------------------ pfilter iteration, ref :298-336
__global__ void k_pf_cut(PmcBufs B, unsigned int N, double q) {
    // Statistics.quantile type 7 on the sorted costs (same operations as the smc quantile)
    const double aleph = xadd(xmul((double)N, q), xsub(1.0, q));
    long long j = (long long)aleph;
    if (j < 1) j = 1;
    if (j > (long long)N - 1) j = (long long)N - 1;
    double gam = xsub(aleph, (double)j);
    gam = gam < 0.0 ? 0.0 : (gam > 1.0 ? 1.0 : gam);
    double a, b;
    if (N == 1) { a = b = pmc_unkey(B.keys[0]); }
    else { a = pmc_unkey(B.keys[j - 1]); b = pmc_unkey(B.keys[j]); }
    const double eps = (dfinite(a) && dfinite(b)) ? xadd(a, xmul(gam, xsub(b, a))) : xadd(xmul(xsub(1.0, gam), a), xmul(gam, b));
    B.ctrl->eps = eps;
    B.ctrl->n_ok = pmc_upper_bound(B.keys, N, pmc_key(eps)); // particles with cost <= eps: the head of the sorted order
}
__global__ void __launch_bounds__(256) k_pf_propose(PmcBufs B, long long N, DPriors pri, RoundKeys rk, const unsigned int *pend,
                                                    unsigned int n_pend, unsigned int *pend2, unsigned int n_ok, uint32_t round,
                                                    double width) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_pend) return;
    const long long i = pend[w];
    Stream st(rk, ST_PROPOSE, (uint32_t)i, round);
    unsigned int b = index_of(st.next(), n_ok), c = b, e;
    while (c == b) c = index_of(st.next(), n_ok);
    e = b;
    while (e == b || e == c) e = index_of(st.next(), n_ok);
    const long long pb = B.sidx[b], pc = B.sidx[c], pe = B.sidx[e];
    const double sc = xmul(next_normal(st), width); // ref :312
    for (int k = 0; k < pri.d; ++k) {
        const double *t = B.th + (long long)k * N;
        B.thp[(long long)k * N + i] = xadd(t[pb], xmul(xsub(t[pe], t[pc]), sc));
    }
    const double *thp = B.thp;
    const double ll = prior_logpdf_pushed(pri, [&](int k) { return thp[(long long)k * N + i]; });
    if (xlog(u01(st.next())) > fmin(0.0, xsub(ll, B.lp[i]))) { // ref :316-318
        pend2[atomicAdd(&B.ctrl->n_pend2, 1u)] = (unsigned int)i;
        return;
    }
    B.lpp[i] = ll;
    B.list[atomicAdd(&B.ctrl->n_eval, 1u)] = (unsigned int)i;
}
__global__ void k_pf_accept(PmcBufs B, long long N, int d, unsigned int *pend2) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int n = B.ctrl->n_eval;
    if (w == 0) B.ctrl->evals += n;
    if (w >= n) return;
    const long long i = B.list[w];
    const double Cp = B.Cp[i];
    if (Cp > B.ctrl->eps) { // ref :320-322 (a NaN cost compares false and is accepted, as in the reference)
        pend2[atomicAdd(&B.ctrl->n_pend2, 1u)] = (unsigned int)i;
        return;
    }
    B.C[i] = Cp;
    for (int k = 0; k < d; ++k) B.th[(long long)k * N + i] = B.thp[(long long)k * N + i];
    B.lp[i] = B.lpp[i];
}

} // namespace kabc

using namespace kabc;

namespace {

constexpr int PMC_INIT_TRIES = 1000;
constexpr int PF_MAX_ROUNDS = 100000;

struct Pmc {
    kabc_ctx *ctx = nullptr;
    DPriors pri;
    DModel model;
    long long N = 0;
    unsigned int M = 0; // N padded to a power of two (sort)
    int d = 0;
    PmcBufs B;
    DevBuf<double> th, lp, C, thp, lpp, Cp, thr;
    DevBuf<unsigned long long> keys;
    DevBuf<unsigned int> sidx, list, pendA, pendB;
    DevBuf<PmcCtrl> ctrl;
    PmcCtrl h;
};

#define PMC_LAUNCHED(p, n) ((p).ctx->launches += (n))

int pmc_setup(Pmc &p, kabc_ctx *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model, long long N) {
    if (!ctx) return set_error(KABC_ERR_INVALID_ARG, "ctx is NULL");
    if (int rc = ingest_priors(prior, d, p.pri)) return rc;
    if (int rc = ingest_model(model, d, p.model)) return rc;
    if (N < 3 || N > 0x7FFFFFFFll) return set_error(KABC_ERR_INVALID_ARG, "nparticles must be in 3..2^31-1");
    p.model.push_mask = 0; // ref :289,:320,:408: the cost sees the raw particle, only the prior the push_p'ed one
    p.ctx = ctx; p.N = N; p.d = d;
    p.M = 1;
    while (p.M < (unsigned long long)N) p.M <<= 1;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t n = (size_t)N;
#define A(x) KABC_CUDA_TRY(x)
    A(p.th.alloc(ctx, n * d)); A(p.lp.alloc(ctx, n)); A(p.C.alloc(ctx, n)); A(p.thp.alloc(ctx, n * d)); A(p.lpp.alloc(ctx, n));
    A(p.Cp.alloc(ctx, n)); A(p.thr.alloc(ctx, n)); A(p.keys.alloc(ctx, p.M)); A(p.sidx.alloc(ctx, p.M)); A(p.list.alloc(ctx, n));
    A(p.pendA.alloc(ctx, n)); A(p.pendB.alloc(ctx, n)); A(p.ctrl.alloc(ctx, 1));
#undef A
    p.B.th = p.th.p; p.B.lp = p.lp.p; p.B.C = p.C.p; p.B.thp = p.thp.p; p.B.lpp = p.lpp.p; p.B.Cp = p.Cp.p; p.B.thr = p.thr.p;
    p.B.keys = p.keys.p; p.B.sidx = p.sidx.p; p.B.list = p.list.p; p.B.pendA = p.pendA.p; p.B.pendB = p.pendB.p;
    p.B.ctrl = p.ctrl.p;
    KABC_CUDA_TRY(cudaMemsetAsync(p.ctrl.p, 0, sizeof(PmcCtrl), ctx->stream));
    return KABC_OK;
}

int pmc_read_ctrl(Pmc &p) {
    KABC_CUDA_TRY(cudaMemcpyAsync(&p.h, p.ctrl.p, sizeof(PmcCtrl), cudaMemcpyDeviceToHost, p.ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(p.ctx->stream));
    if (p.h.err) return set_error(p.h.err, "prior sampling failed (truncation too extreme)");
    return KABC_OK;
}

unsigned blocks_for(long long n) { return (unsigned)((n + 255) / 256); }

// particles drawn until cost and log-prior are finite (one round per retry over the still-pending ones)
int pmc_init(Pmc &p) {
    cudaStream_t st = p.ctx->stream;
    k_pmc_iota<<<blocks_for(p.N), 256, 0, st>>>(p.pendA.p, (unsigned int)p.N);
    PMC_LAUNCHED(p, 1);
    unsigned int *pend = p.pendA.p, *pend2 = p.pendB.p;
    unsigned int n_pend = (unsigned int)p.N;
    for (int t = 0; n_pend > 0; ++t) {
        if (t >= PMC_INIT_TRIES) return set_error(KABC_ERR_RETRY_BUDGET, "Prior leads to non-finite costs too often");
        k_pmc_reset_round<<<1, 1, 0, st>>>(p.ctrl.p);
        k_pmc_init_propose<<<blocks_for(n_pend), 256, 0, st>>>(p.B, p.N, p.pri, p.ctx->rk, pend, n_pend, pend2, (uint32_t)t);
        PMC_LAUNCHED(p, 2);
        if (int rc = eval_cost_list_device(p.ctx, p.model, p.th.p, p.N, p.list.p, &p.ctrl.p->n_eval, n_pend, ST_COST_INIT, (uint32_t)t, p.Cp.p))
            return rc;
        k_pmc_init_accept<<<blocks_for(n_pend), 256, 0, st>>>(p.B, pend2);
        PMC_LAUNCHED(p, 1);
        KABC_CUDA_TRY(cudaGetLastError());
        if (int rc = pmc_read_ctrl(p)) return rc;
        n_pend = p.h.n_pend2;
        unsigned int *tmp = pend; pend = pend2; pend2 = tmp;
    }
    return KABC_OK;
}

int pmc_sort(Pmc &p) {
    cudaStream_t st = p.ctx->stream;
    k_pmc_sort_fill<<<blocks_for(p.M), 256, 0, st>>>(p.C.p, (unsigned int)p.N, p.M, p.keys.p, p.sidx.p);
    PMC_LAUNCHED(p, 1);
    if (p.M <= (unsigned)PMC_SORT_BLOCK) {
        k_pmc_sort_block<<<1, 1024, 0, st>>>(p.keys.p, p.sidx.p, p.M);
        PMC_LAUNCHED(p, 1);
    } else {
        for (unsigned int k = 2; k <= p.M; k <<= 1)
            for (unsigned int j = k >> 1; j > 0; j >>= 1) {
                k_pmc_sort_step<<<blocks_for(p.M), 256, 0, st>>>(p.keys.p, p.sidx.p, p.M, j, k);
                PMC_LAUNCHED(p, 1);
            }
    }
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

// host copies of the final state; discrete components are returned push_p'ed (ref :338, :421)
int pmc_output(Pmc &p, double *out_theta, double *out_cost) {
    cudaStream_t st = p.ctx->stream;
    const size_t n = (size_t)p.N;
    if (out_theta) KABC_CUDA_TRY(cudaMemcpyAsync(out_theta, p.th.p, sizeof(double) * n * p.d, cudaMemcpyDeviceToHost, st));
    if (out_cost) KABC_CUDA_TRY(cudaMemcpyAsync(out_cost, p.C.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    if (out_theta)
        for (int k = 0; k < p.d; ++k)
            if (prior_is_discrete(p.pri.p[k]))
                for (size_t i = 0; i < n; ++i) out_theta[(size_t)k * n + i] = nearbyint(out_theta[(size_t)k * n + i]);
    return KABC_OK;
}

} // namespace

extern "C" {

int kabc_abcde_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model, const kabc_abcde_config_t *cfg,
                   double *out_theta, double *out_cost, int32_t *out_reached, int64_t *out_nsim, int64_t *out_generations) {
    if (!cfg) return set_error(KABC_ERR_INVALID_ARG, "cfg is NULL");
    if (!(cfg->alpha >= 0.0 && cfg->alpha < 1.0)) return set_error(KABC_ERR_INVALID_ARG, "α must be in 0 <= α < 1."); // ref :353
    if (cfg->generations < 0) return set_error(KABC_ERR_INVALID_ARG, "generations must be >= 0");
    Pmc p;
    if (int rc = pmc_setup(p, ctx, prior, d, model, cfg->nparticles)) return rc;
    if (int rc = pmc_init(p)) return rc;
    cudaStream_t st = ctx->stream;
    KABC_CUDA_TRY(cudaMemsetAsync(&p.ctrl.p->evals, 0, sizeof(unsigned long long), st)); // ref :373 nsims starts after the init
    AbcdeParams A;
    A.eps_target = cfg->eps_target; A.alpha = cfg->alpha; A.earlystop = cfg->earlystop ? 1 : 0;
    A.gamma = (cfg->proposal_width * 2.38) / sqrt((double)(2 * d)); // ref :374
    for (long long g = 1; g <= cfg->generations; ++g) {
        if (int rc = pmc_sort(p)) return rc;
        k_abcde_pre<<<1, 1, 0, st>>>(p.B, (unsigned int)p.N, A);
        k_abcde_propose<<<blocks_for(p.N), 256, 0, st>>>(p.B, p.N, p.pri, ctx->rk, A);
        PMC_LAUNCHED(p, 2);
        if (int rc = eval_cost_list_device(ctx, p.model, p.thp.p, p.N, p.list.p, &p.ctrl.p->n_eval, p.N, ST_COST, (uint32_t)g, p.Cp.p))
            return rc;
        k_abcde_accept<<<blocks_for(p.N), 256, 0, st>>>(p.B, p.N, d);
        PMC_LAUNCHED(p, 1);
        KABC_CUDA_TRY(cudaGetLastError());
    }
    if (int rc = pmc_sort(p)) return rc; // for the final maximum
    if (int rc = pmc_read_ctrl(p)) return rc;
    unsigned long long kmax = 0;
    KABC_CUDA_TRY(cudaMemcpyAsync(&kmax, p.keys.p + (p.N - 1), 8, cudaMemcpyDeviceToHost, st));
    if (int rc = pmc_output(p, out_theta, out_cost)) return rc;
    const unsigned long long bits = (kmax >> 63) ? (kmax & 0x7FFFFFFFFFFFFFFFull) : ~kmax;
    double cmax;
    memcpy(&cmax, &bits, 8);
    if (out_reached) *out_reached = cmax <= cfg->eps_target; // ref :418
    if (out_nsim) *out_nsim = (int64_t)p.h.evals;
    if (out_generations) *out_generations = p.h.gens;
    return KABC_OK;
}

int64_t kabc_pfilter_nparticles(int64_t n, int d, double q) { // ref :276-279
    const int64_t lowN = 4 * (int64_t)d;
    if ((double)n * q <= (double)lowN) n = (int64_t)ceil((double)(lowN + 1) / q);
    return n;
}

int kabc_pfilter_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model, const kabc_pfilter_config_t *cfg,
                     double *out_theta, double *out_cost, double *out_eps, int64_t *out_iterations, int64_t *out_nreps,
                     int64_t *out_cost_evals) {
    if (!cfg) return set_error(KABC_ERR_INVALID_ARG, "cfg is NULL");
    if (!(cfg->q > 0.0 && cfg->q <= 1.0)) return set_error(KABC_ERR_INVALID_ARG, "pfilter needs 0 < q <= 1");
    if (d < 1 || d > KABC_MAX_DIM) return set_error(KABC_ERR_INVALID_ARG, "d must be in 1..%d", KABC_MAX_DIM);
    Pmc p;
    if (int rc = pmc_setup(p, ctx, prior, d, model, kabc_pfilter_nparticles(cfg->nparticles, d, cfg->q))) return rc;
    if (int rc = pmc_init(p)) return rc;
    cudaStream_t st = ctx->stream;
    long long iters = 0, reps_total = 0;
    uint32_t round = 0; // epoch of the attempt streams: one per rejection round over the whole run
    for (;;) {
        iters += 1;
        if (int rc = pmc_sort(p)) return rc;
        k_pf_cut<<<1, 1, 0, st>>>(p.B, (unsigned int)p.N, cfg->q);
        PMC_LAUNCHED(p, 1);
        if (int rc = pmc_read_ctrl(p)) return rc;
        const unsigned int n_ok = p.h.n_ok, n_bad = (unsigned int)p.N - n_ok;
        if (n_bad > 0 && n_ok < 3) return set_error(KABC_ERR_DEGENERATE, "pfilter: fewer than 3 particles under the quantile");
        // the bad particles are the tail of the sorted order
        const unsigned int *pend = p.sidx.p + n_ok;
        unsigned int *pend2 = p.pendA.p, *spare = p.pendB.p;
        unsigned int n_pend = n_bad;
        long long nreps = 0;
        for (int r = 0; n_pend > 0; ++r) {
            if (r >= PF_MAX_ROUNDS) return set_error(KABC_ERR_RETRY_BUDGET, "pfilter: rejection loop does not terminate");
            round += 1;
            nreps += n_pend; // ref :313: every attempt counts
            k_pmc_reset_round<<<1, 1, 0, st>>>(p.ctrl.p);
            k_pf_propose<<<blocks_for(n_pend), 256, 0, st>>>(p.B, p.N, p.pri, ctx->rk, pend, n_pend, pend2, n_ok, round, cfg->proposal_width);
            PMC_LAUNCHED(p, 2);
            if (int rc = eval_cost_list_device(ctx, p.model, p.thp.p, p.N, p.list.p, &p.ctrl.p->n_eval, n_pend, ST_COST, round, p.Cp.p))
                return rc;
            k_pf_accept<<<blocks_for(n_pend), 256, 0, st>>>(p.B, p.N, d, pend2);
            PMC_LAUNCHED(p, 1);
            KABC_CUDA_TRY(cudaGetLastError());
            if (int rc = pmc_read_ctrl(p)) return rc;
            n_pend = p.h.n_pend2;
            pend = pend2;
            unsigned int *tmp = pend2; pend2 = spare; spare = tmp;
        }
        reps_total += nreps;
        // ref :331-335.  An iteration without bad particles (eff = 0/0 in the reference, which then never leaves its
        // loop) ends the run.
        if (n_bad == 0) break;
        const double eff = (double)n_bad / (double)nreps;
        if (eff < cfg->eff_tol) break;
        if (p.h.eps < cfg->epstol) break;
        if (cfg->max_iters > 0 && iters > cfg->max_iters) break;
    }
    if (int rc = pmc_output(p, out_theta, out_cost)) return rc;
    if (out_eps) *out_eps = p.h.eps;
    if (out_iterations) *out_iterations = iters;
    if (out_nreps) *out_nreps = reps_total;
    if (out_cost_evals) *out_cost_evals = (int64_t)p.h.evals;
    return KABC_OK;
}

} // extern "C"
