/*
 * kissabc_cuda.h -- C ABI of libkissabc_cuda.so, the B200 (sm_100a) implementation of the
 * KissABC.jl hot path: the per-particle propose -> simulate -> distance -> accept loop behind
 *     sample(ApproxKernelizedPosterior(prior, cost, eps), AIS(N), Ns; ...)   and   smc(prior, cost; ...)
 *
 * The reference (KissABC.jl 3.0.1, pure Julia) has NO FFI / plugin interface; its only seams are
 * Julia-level:
 *   - the `cost` callable                     src/types.jl:55, src/smc.jl:123,176
 *   - the density protocol                    src/types.jl:3-8 (unconditional_sample, loglike, accept, push_p)
 *   - the AbstractMCMC sampler protocol       src/KissABC.jl:35-80 (two `step` methods), :82-94 (bundle_samples)
 *   - `smc(prior, cost; kw...)`               src/smc.jl:92-106
 * Each entry point below names the reference code it replaces.  A Julia maintainer binds them with
 * `ccall` (INTEGRATION.md shows the stub); this repo's tested host mirror binds the same symbols with ctypes.
 *
 * Conventions: every function returns an int status (0 = KABC_OK); kabc_last_error() returns the message
 * of the last failure on the calling thread.  The caller owns all host buffers; the library owns device
 * memory behind opaque handles.  One handle = one host thread at a time.  No callbacks into the host.
 * All floating-point state is FP64 (the reference is Float64 end to end, src/smc.jl:119).  Arrays of
 * parameters are SoA, "k-major": theta[k*n + i] is parameter k of particle i (a Julia n x d column-major
 * matrix).  Particle indices at the boundary are 0-based.
 */
#ifndef KISSABC_CUDA_H
#define KISSABC_CUDA_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KABC_VERSION 100 /* 0.1.0 */

enum kabc_status {
    KABC_OK = 0,
    KABC_ERR_INVALID_ARG = 1,  /* the reference's error(...) on bad arguments: src/smc.jl:107-118, src/KissABC.jl:43-48 */
    KABC_ERR_CUDA = 2,
    KABC_ERR_NCCL = 3,
    KABC_ERR_RETRY_BUDGET = 4, /* "Prior leads to inf costs too often", src/KissABC.jl:58-59 */
    KABC_ERR_DEGENERATE = 5,   /* no alive particles left to resample (Julia would throw at src/smc.jl:147) */
    KABC_ERR_STATE = 6,        /* handle used out of order; also the reference's "starting sample invalid" / "ld_correction is
                                  invalid" errors of accept (src/types.jl:69-70) */
    KABC_ERR_PEER = 7          /* a rank of the job did not reach a cross-rank barrier in time (peer memory, multi GPU) */
};

/* ---- prior: Factored(Uniform(a,b), Normal(mu,sigma), Truncated(Normal(mu,sigma),lo,hi), Beta(a,b),
 *      NegativeBinomial(r,p), DiscreteUniform(a,b), ...)
 *      replaces src/priors.jl:10-49 + the Distributions.jl univariate laws it delegates to.  Components of a discrete law
 *      (NegativeBinomial, DiscreteUniform) follow push_p, src/types.jl:28-32: the particle keeps its real value, the prior
 *      density and the cost see round(Int, .), and so do the returned samples ---- */
enum kabc_prior_kind {
    KABC_PRIOR_UNIFORM = 0,
    KABC_PRIOR_NORMAL = 1,
    KABC_PRIOR_TRUNC_NORMAL = 2,
    KABC_PRIOR_BETA = 3,            /* test/runtests.jl:51, examples/example_n2.jl:28 */
    KABC_PRIOR_NEG_BINOMIAL = 4,    /* test/runtests.jl:50 */
    KABC_PRIOR_DISCRETE_UNIFORM = 5 /* test/runtests.jl:106 */
};
typedef struct {
    int32_t kind;
    int32_t _pad;
    double p0, p1; /* Uniform: a,b.  Normal / Truncated(Normal): mu, sigma.  Beta: alpha, beta.  NegativeBinomial: r, p.
                      DiscreteUniform: a, b (integers) */
    double lo, hi; /* Truncated: support bounds (ignored otherwise) */
} kabc_prior_t;
#define KABC_MAX_DIM 16

/* ---- model: the registered device simulator + distance that replaces the opaque `cost` closure
 *      (src/types.jl:55, src/smc.jl:123,176).  Definitions: DESIGN.md "Simulators" ---- */
enum kabc_model_kind {
    KABC_MODEL_NORMAL_MEANSTD = 0, /* README.md:35-52: x=randn(n)*sigma+mu; hypot(mean-t0,(std-t1)*param0) */
    KABC_MODEL_MA2_AUTOCOV = 1,    /* MA(2), n obs; || (tau1,tau2) - target ||_2; +Inf outside the triangle */
    KABC_MODEL_GK_OCTILE = 2,      /* g-and-k, n draws, 7 octiles; param0 = c (0.8) */
    KABC_MODEL_LV_SSA = 3,         /* Lotka-Volterra Gillespie; param = {X0,Y0,T,G,max_events}; target = [X(t_g), Y(t_g)] */
    KABC_MODEL_DETERMINISTIC = 4,  /* test/runtests.jl:77-86 (param0=0: |th^2+1-t0|), :177-182 (param0=1: |th-t0|) and
                                      :105-112 (param0=2, d=2: |(th0^2+th1)*(th0+randn*param1) - t0|) */
    KABC_MODEL_SOCKS = 5           /* test/runtests.jl:34-44, d=2 (n_socks, prop_pairs); param0 = n_picked (<= 32);
                                      cost = |pairs - t0| + |odds - t1| */
};
enum kabc_precision {
    KABC_F64 = 0,      /* simulator entirely in FP64, bit-reproducible against the oracle */
    KABC_F32_ACC64 = 1 /* simulator draws in FP32 (MUFU Box-Muller), distance finished in FP64 */
};
#define KABC_MAX_TARGET 32
#define KABC_MAX_PARAM 8
typedef struct {
    int32_t kind;
    int32_t precision;
    int32_t n_draws;
    int32_t n_target;
    double target[KABC_MAX_TARGET];
    double param[KABC_MAX_PARAM];
} kabc_model_t;

/* ---- smc keyword arguments: same names and defaults as src/smc.jl:95-105 (`parallel` has no meaning
 *      on the device; `rng` is replaced by the context seed) ---- */
typedef struct {
    int64_t nparticles;    /* 100 */
    double alpha;          /* 0.95 */
    int64_t mcmc_retrys;   /* 0 */
    double mcmc_tol;       /* 0.015 */
    double epstol;         /* 0.0 */
    double r_epstol;       /* (1-alpha)^1.5/50 */
    double min_r_ess;      /* alpha^2 */
    double max_stretch;    /* 2.0 */
    int32_t verbose;       /* 0 */
    int32_t max_iterations; /* 0 = unbounded (reference behaviour); extension used by benchmarks */
} kabc_smc_config_t;

/* ---- AIS arguments: AIS(nwalkers), sample(..., nsamples; ntransitions, discard_initial, thinning),
 *      retry_sampling (src/KissABC.jl:39), scale = target_average_cost of ApproxKernelizedPosterior
 *      (src/types.jl:40-49).  Move mixture 4/7,2/7,1/7 and stretch a = 3 are fixed (src/transition.jl:56,62) ---- */
typedef struct {
    int64_t nwalkers;
    int64_t nsamples;
    int64_t ntransitions;    /* 1 */
    int64_t discard_initial; /* 0 */
    int64_t thinning;        /* 1 */
    int64_t retry_sampling;  /* 100 */
    double scale;            /* target_average_cost (posterior 0) or max_cost (posterior 1) */
    int32_t posterior;       /* 0 = ApproxKernelizedPosterior (src/types.jl:40-75), 1 = ApproxPosterior (src/types.jl:76-104) */
    int32_t _pad;
} kabc_ais_config_t;

/* one record per smc iteration: what `verbose && @show iteration, eps, ESS` prints (src/smc.jl:143) + counters */
typedef struct {
    int64_t iteration;
    double eps;
    int64_t n_alive;    /* ESS after the cut, before resampling */
    int32_t flag;       /* src/smc.jl:135-141 */
    int32_t resampled;  /* src/smc.jl:145 */
    int64_t accepted;   /* src/smc.jl:156,186 */
    int64_t cost_evals; /* cumulative, including the nparticles evaluations of the initialisation */
    int64_t sweeps;     /* MCMC sweeps executed in this iteration (<= 1+mcmc_retrys) */
} kabc_smc_log_t;

typedef struct kabc_ctx kabc_ctx_t;
typedef struct kabc_smc kabc_smc_t;
typedef struct kabc_ais kabc_ais_t;

/* ---- introspection ---- */
int kabc_version(void);
const char *kabc_last_error(void);
int kabc_device_count(int *count);

/* ---- context: device + stream + Philox seed (+ NCCL communicator when world > 1).
 *      Replaces the `rng` argument of smc (src/smc.jl:95) and sample. ---- */
#define KABC_NCCL_ID_BYTES 128
int kabc_nccl_unique_id(char id[KABC_NCCL_ID_BYTES]); /* rank 0 creates, the host distributes it */
int kabc_ctx_create(int device, uint64_t seed, kabc_ctx_t **ctx);
int kabc_ctx_create_dist(int device, uint64_t seed, int rank, int world, const char id[KABC_NCCL_ID_BYTES],
                         kabc_ctx_t **ctx);
/* Multi-rank jobs (one process per GPU) synchronise and exchange rows through a PEER ARENA: one device allocation per rank,
 * mapped into every other rank with cudaIpc (NVLink peer memory) and kept for the life of the context.  A context made by
 * kabc_ctx_create_dist sizes, exchanges and maps its arena by itself (the 64-byte handles travel through its NCCL
 * communicator; NCCL is not used on the data path).  kabc_ctx_create_ranks makes a rank of a job WITHOUT NCCL -- e.g. several
 * ranks sharing one GPU, which NCCL refuses -- and leaves the exchange to the host: every rank calls
 * kabc_ctx_arena_export(bytes), the host all-gathers the handles (MPI, torch.distributed, Distributed.jl, ...) and every rank
 * calls kabc_ctx_arena_attach with the world x 64 bytes in rank order.  bytes: kabc_smc_arena_bytes / kabc_ais_arena_bytes. */
#define KABC_IPC_HANDLE_BYTES 64
int kabc_ctx_create_ranks(int device, uint64_t seed, int rank, int world, kabc_ctx_t **ctx);
int kabc_ctx_arena_export(kabc_ctx_t *ctx, uint64_t bytes, char handle[KABC_IPC_HANDLE_BYTES]);
int kabc_ctx_arena_attach(kabc_ctx_t *ctx, const char *handles /* world x KABC_IPC_HANDLE_BYTES */);
uint64_t kabc_smc_arena_bytes(int64_t nparticles, int d, int world);
uint64_t kabc_ais_arena_bytes(int64_t nwalkers, int d, int world);
int kabc_ctx_destroy(kabc_ctx_t *ctx); /* KABC_ERR_STATE while smc / ais handles of the context are alive */
int kabc_ctx_info(const kabc_ctx_t *ctx, int *device, int *rank, int *world, int *sm_count);

/* ---- priors on device: logpdf(Factored, x) src/priors.jl:30-36; rand(rng, Factored) src/priors.jl:42-43 ---- */
int kabc_prior_logpdf(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const double *theta, int64_t n,
                      double *out_logpdf);
int kabc_prior_sample(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, int64_t n, uint32_t first_id,
                      uint32_t epoch, double *out_theta);

/* ---- bare cost evaluation: `cost(x)` for n parameter vectors (src/types.jl:55, src/smc.jl:123,176).
 *      Particle i uses the COST stream (first_id+i, epoch).  out_events may be NULL (LV event counts). ---- */
int kabc_eval_cost(kabc_ctx_t *ctx, const kabc_model_t *model, int d, const double *theta, int64_t n,
                   uint32_t first_id, uint32_t epoch, double *out_cost, int64_t *out_events);
/* same with theta / out already resident in device memory (benchmark + composition) */
int kabc_eval_cost_device(kabc_ctx_t *ctx, const kabc_model_t *model, int d, const double *d_theta, int64_t n,
                          uint32_t first_id, uint32_t epoch, double *d_out_cost, float *out_ms);

/* ---- smc(prior, cost; kw...) -> (P, C, eps): src/smc.jl:92-206 in one call.
 *      out_theta[d*N] (push_p'd), out_alive[N] (P = theta[:, alive]), out_cost[N] (C = Xs, all N) ---- */
int kabc_smc_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                 const kabc_smc_config_t *cfg, double *out_theta, uint8_t *out_alive, double *out_cost,
                 double *out_eps, int64_t *out_iterations, int64_t *out_cost_evals, kabc_smc_log_t *log,
                 int64_t log_cap);

/* ---- the same, step by step (state stays resident in HBM between calls) ---- */
int kabc_smc_create(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                    const kabc_smc_config_t *cfg, kabc_smc_t **smc);
int kabc_smc_destroy(kabc_smc_t *smc);
int kabc_smc_init(kabc_smc_t *smc);                /* src/smc.jl:119-129 */
int kabc_smc_iterate(kabc_smc_t *smc, int *stop);  /* one body of `while true`, src/smc.jl:131-198;
                                                      stop: 0 continue, 1 r_epstol, 2 epstol, 3 acceptance, 4 max_iterations */
int kabc_smc_iterate_n(kabc_smc_t *smc, int n, int ignore_stop, int *done, float *out_ms); /* n bodies back to back */
int kabc_smc_get_state(kabc_smc_t *smc, double *theta, double *X, double *lpi, uint8_t *alive);
int kabc_smc_set_state(kabc_smc_t *smc, const double *theta, const double *X, const double *lpi, const uint8_t *alive);
int kabc_smc_get_scalars(kabc_smc_t *smc, double *eps, int32_t *flag, int64_t *iteration, int64_t *n_alive,
                         int64_t *accepted, int64_t *cost_evals, int64_t *next_epoch, int64_t *events);
int64_t kabc_smc_get_log(kabc_smc_t *smc, kabc_smc_log_t *log, int64_t cap);
int64_t kabc_smc_kernel_launches(kabc_smc_t *smc);
/* benchmark helper: n bodies enqueued back to back (stop rules ignored), each preceded by a flush of the L2 cache
 * (a device memset of flush_bytes, outside the timed region); out_ms[i] = device time of body i between two events */
int kabc_smc_bench_steps(kabc_smc_t *smc, int n, uint64_t flush_bytes, float *out_ms);
/* one iteration with CUDA events between its kernels: warm per-kernel times in microseconds, in launch order
 * (select x3, cut, table, sweep -- or table, propose, simulate for the work-list simulators) */
int kabc_smc_profile_iteration(kabc_smc_t *smc, float *out_us, int cap, int *out_n);
/* replay hooks: record, for the next sweeps, the variates and decisions each particle used so that the CPU
 * oracle can replay them through the reference logic (north_star "Philox uniforms replayed").
 * decision: 0 dead, 1 prior -Inf, 2 failed prior-MH pre-test, 3 simulated+rejected, 4 accepted */
int kabc_smc_trace_enable(kabc_smc_t *smc, int on);
int kabc_smc_get_trace(kabc_smc_t *smc, int64_t *a, int64_t *b, double *z, double *lprob, double *lpi_p,
                       double *xp, uint8_t *decision, double *theta_p);

/* ---- sample(ApproxKernelizedPosterior(prior,cost,scale), AIS(N), Ns; ...): src/KissABC.jl:35-94,
 *      src/transition.jl:2-82, src/types.jl:51-75.  out_samples[d*nsamples], SoA ---- */
int kabc_ais_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                 const kabc_ais_config_t *cfg, double *out_samples, int64_t *out_cost_evals, int64_t *out_accepted);
int kabc_ais_create(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                    const kabc_ais_config_t *cfg, kabc_ais_t **ais);
int kabc_ais_destroy(kabc_ais_t *ais);
int kabc_ais_init(kabc_ais_t *ais);                            /* src/KissABC.jl:50-61 */
int kabc_ais_sweep(kabc_ais_t *ais, int nsweeps, float *out_ms); /* red/black sweeps: 2 half-steps each */
int kabc_ais_get_state(kabc_ais_t *ais, double *theta, double *logprior, double *loglike);
int kabc_ais_set_state(kabc_ais_t *ais, const double *theta, const double *logprior, const double *loglike);
int kabc_ais_get_counters(kabc_ais_t *ais, int64_t *cost_evals, int64_t *accepted, int64_t *sweeps, int64_t *retries);
int64_t kabc_ais_kernel_launches(kabc_ais_t *ais);
/* decision: 0 new state invalid (no variate consumed), 1 rejected, 2 accepted; move 1 stretch, 2 DE, 3 walk */
int kabc_ais_trace_enable(kabc_ais_t *ais, int on);
int kabc_ais_get_trace(kabc_ais_t *ais, uint8_t *move, int64_t *a, int64_t *b, int64_t *c, double *corr,
                       double *theta_p, double *lp_p, double *ll_p, double *e, uint8_t *decision);

/* ---- ABCDE(prior, cost, eps_target; nparticles, generations, alpha, earlystop, proposal_width): src/smc.jl:352-428.
 *      Population Monte Carlo with differential-evolution moves; every generation proposes for all particles from the
 *      previous generation's state (Jacobi update, ref :379-381).  out_theta: d x nparticles SoA, push_p'ed (ref :421);
 *      out_cost: nparticles; out_reached: maximum(cost) <= eps_target (ref :418); out_nsim: sum(nsims) (ref :407) ---- */
typedef struct {
    int64_t nparticles;    /* 50 */
    int64_t generations;   /* 20 */
    double eps_target;
    double alpha;          /* 0 <= alpha < 1 (ref :353) */
    double proposal_width; /* 1.0 */
    int32_t earlystop;     /* false */
    int32_t _pad;
} kabc_abcde_config_t;
int kabc_abcde_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model, const kabc_abcde_config_t *cfg,
                   double *out_theta, double *out_cost, int32_t *out_reached, int64_t *out_nsim, int64_t *out_generations);

/* ---- pfilter(prior, cost, N; q, eff_tol, epstol, max_iters, proposal_width): src/smc.jl:275-345.
 *      Every iteration cuts at the q-quantile of the costs and redraws each particle above it from three particles
 *      below it until prior pre-test and cost pass.  The particle count is kabc_pfilter_nparticles(N, d, q)
 *      (ref :276-279); the caller sizes out_theta (d x that, SoA, push_p'ed) and out_cost with it.
 *      max_iters = 0 means Inf.  Deviation: an iteration with no particle above the quantile ends the run (the
 *      reference computes eff = 0/0 there and loops forever) ---- */
typedef struct {
    int64_t nparticles;
    double q;              /* 0.7 */
    double eff_tol;        /* 0.1 */
    double epstol;         /* -Inf */
    double proposal_width; /* 0.75 */
    int64_t max_iters;     /* 0 = Inf */
} kabc_pfilter_config_t;
int64_t kabc_pfilter_nparticles(int64_t n, int d, double q);
int kabc_pfilter_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model, const kabc_pfilter_config_t *cfg,
                     double *out_theta, double *out_cost, double *out_eps, int64_t *out_iterations, int64_t *out_nreps,
                     int64_t *out_cost_evals);

/* ---- instruction-pipe microbenchmarks used to state the issue roofline (DESIGN.md "Roofline") ----
 * kind: 0 FFMA, 1 IMAD, 2 IMAD.WIDE(mul.wide.u32), 3 LOP3, 4 MUFU.LG2, 5 MUFU.SIN, 6 MUFU.SQRT, 7 DFMA,
 *       8 I2F, 9 Philox4x32-10 words, 10 F32 Box-Muller normals, 11 F64 spec normals.
 * out_rate = thread-level operations per second over the whole GPU. */
int kabc_microbench(kabc_ctx_t *ctx, int kind, double *out_rate, float *out_ms);

#ifdef __cplusplus
}
#endif
#endif
