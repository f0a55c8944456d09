/*
 * kabc_oracle.h -- CPU ORACLE for the KissABC hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference algorithm
 * (/root/reference, KissABC.jl 3.0.1):
 *     src/transition.jl:2-82   (stretch / DE / walk proposals, transition!)
 *     src/types.jl:10-32,51-75 (op fold order, push_p, kernelized loglike, accept)
 *     src/priors.jl:30-43      (Factored logpdf / rand)
 *     src/KissABC.jl:35-80     (AIS init+retry, step, walker rotation)
 *     src/smc.jl:92-206        (smc)
 *     README.md:35-52          (normal model cost)
 * over the canonical counter-based variate source described in DESIGN.md
 * ("Variate spec").  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.  The CUDA library never links or calls it.
 *
 * PARITY STATUS: the reference cannot be executed in this environment (no
 * Julia), and its tests hold no stream-level golden vectors, so *stream-level
 * parity is unpinned*.  What IS pinned: the exact Factored/Uniform values of
 * test/runtests.jl:8-22, the type-7 quantile against numpy, the README
 * posterior (README.md:64-66,84) and the statistical fixtures of
 * test/runtests.jl:77-86,133-175 (see tests/test_oracle_*.py).
 * Cross-check (not a reference run): the SERIAL mode below takes the same decisions as a line-by-line Python transliteration
 * of the reference's Julia on seven whole runs (tests/golden/make_pyref_fixtures.py, tests/test_ref_fixtures.py); the same
 * test consumes tests/golden/ref_*.json once julia/make_ref_fixtures.jl has been run on a box with Julia.
 */
#ifndef KABC_ORACLE_H
#define KABC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- descriptors (same field order as include/kissabc_cuda.h PODs) ---- */
enum { KOR_PRIOR_UNIFORM = 0, KOR_PRIOR_NORMAL = 1, KOR_PRIOR_TRUNC_NORMAL = 2,
       KOR_PRIOR_BETA = 3, KOR_PRIOR_NEG_BINOMIAL = 4, KOR_PRIOR_DISCRETE_UNIFORM = 5 };
typedef struct {
    int32_t kind;
    int32_t _pad;
    double p0, p1;   /* Uniform(a,b) | Normal(mu,sigma) | Truncated(Normal(mu,sigma),lo,hi) */
    double lo, hi;
} kor_prior_t;

enum { KOR_MODEL_NORMAL_MEANSTD = 0, KOR_MODEL_MA2_AUTOCOV = 1, KOR_MODEL_GK_OCTILE = 2,
       KOR_MODEL_LV_SSA = 3, KOR_MODEL_DETERMINISTIC = 4, KOR_MODEL_SOCKS = 5 };
#define KOR_MAX_TARGET 32
#define KOR_MAX_PARAM 8
typedef struct {
    int32_t kind;
    int32_t precision;  /* ignored by the oracle: always FP64 */
    int32_t n_draws;
    int32_t n_target;
    double target[KOR_MAX_TARGET];
    double param[KOR_MAX_PARAM];
} kor_model_t;

typedef struct {
    int64_t nparticles;
    double alpha;
    int64_t mcmc_retrys;
    double mcmc_tol;
    double epstol;
    double r_epstol;
    double min_r_ess;
    double max_stretch;
    int32_t verbose;
    int32_t max_iterations; /* 0 = unbounded (reference behaviour) */
} kor_smc_config_t;

typedef struct {
    int64_t nwalkers;
    int64_t nsamples;
    int64_t ntransitions;
    int64_t discard_initial;
    int64_t thinning;
    int64_t retry_sampling;
    double scale;      /* kernel scale (posterior 0) or max_cost (posterior 1) */
    int32_t posterior; /* 0 ApproxKernelizedPosterior, 1 ApproxPosterior (hard threshold) */
    int32_t _pad;
} kor_ais_config_t;

/* per-iteration log record (same layout as kabc_smc_log_t) */
typedef struct {
    int64_t iteration;
    double eps;
    int64_t n_alive;   /* ESS printed by the reference: count after the cut, before resampling */
    int32_t flag;
    int32_t resampled;
    int64_t accepted;
    int64_t cost_evals; /* cumulative, incl. the N of the initialisation */
    int64_t sweeps;     /* MCMC sweeps executed in this iteration */
} kor_smc_log_t;

/* ---- variate spec primitives ---- */
void kor_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double kor_log(double x);
double kor_lgamma(double x); /* x > 0 */
double kor_exp(double x);
void kor_sincos2pi(double u, double *s, double *c);
double kor_u01(uint32_t w);
uint32_t kor_index(uint32_t w, uint32_t n);
void kor_normal_pair(uint32_t w0, uint32_t w1, double *z0, double *z1);
/* word `k` (0-based) of stream (seed, stream, id, epoch) */
uint32_t kor_stream_word(uint64_t seed, uint32_t stream, uint32_t id, uint32_t epoch, uint32_t k);

/* ---- priors: src/priors.jl:30-43 ---- */
double kor_prior_logpdf(const kor_prior_t *prior, int d, const double *x);
int kor_prior_sample(uint64_t seed, const kor_prior_t *prior, int d, uint32_t id, uint32_t epoch, double *x);
/* ref src/types.jl:28-32: discrete components rounded (ties to even), continuous ones untouched; out may alias x */
void kor_push_p(const kor_prior_t *prior, int d, const double *x, double *out);

/* ---- cost: simulator + distance ---- */
double kor_cost(const kor_model_t *model, uint64_t seed, int d, const double *theta, uint32_t id, uint32_t epoch);
/* LV event counter of the last kor_cost call on this thread (roofline unit of config 5) */
int64_t kor_last_events(void);
void kor_eval_cost(const kor_model_t *model, uint64_t seed, int d, const double *theta_soa, int64_t n,
                   uint32_t first_id, uint32_t epoch, double *out, int nthreads);

int kor_lv_trajectory(const kor_model_t *m, uint64_t seed, const double *theta, uint32_t id, uint32_t epoch, double *out);

/* ---- Statistics.quantile type 7 (src/smc.jl:134) on a scratch copy ---- */
double kor_quantile7(const double *v, int64_t n, double p);

/* ---- smc: src/smc.jl:92-206 ---- */
typedef struct kor_smc kor_smc_t;
int kor_smc_create(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model,
                   const kor_smc_config_t *cfg, int nthreads, kor_smc_t **out);
void kor_smc_destroy(kor_smc_t *s);
const char *kor_last_error(void);
int kor_smc_init(kor_smc_t *s);
/* one pass of the `while true` body; *stop = 0 continue, 1 r_epstol, 2 epstol, 3 acceptance, 4 max_iterations */
int kor_smc_iterate(kor_smc_t *s, int *stop);
int kor_smc_run(kor_smc_t *s);
/* kor_smc_iterate in parts, for emulating the sharded multi-rank schedule (tests/test_dist_gloo.py) */
int kor_smc_cut(kor_smc_t *s);
void kor_smc_sweep_range(kor_smc_t *s, int64_t lo, int64_t hi, int64_t *acc, int64_t *evals, int64_t *events);
int kor_smc_sweep_commit(kor_smc_t *s, int64_t acc, int64_t evals, int64_t events);
int kor_smc_finish(kor_smc_t *s, int *stop);
/* replay hook: costs for the NEXT sweeps are taken from xp[i] instead of simulated (NULL = off) */
void kor_smc_set_cost_override(kor_smc_t *s, const double *xp);
/* every variate from ONE word stream (seed; tag 6, id 0, epoch 0) in the reference's own consumption order: what the unmodified
 * KissABC.smc consumes from julia/PhiloxRNG.jl (call before kor_smc_init) */
void kor_smc_set_serial(kor_smc_t *s);
void kor_smc_get_state(const kor_smc_t *s, double *theta_soa, double *X, double *lpi, uint8_t *alive);
void kor_smc_set_state(kor_smc_t *s, const double *theta_soa, const double *X, const double *lpi, const uint8_t *alive);
void kor_smc_get_scalars(const kor_smc_t *s, double *eps, int32_t *flag, int64_t *iteration, int64_t *n_alive,
                         int64_t *accepted, int64_t *cost_evals, int64_t *next_epoch);
int64_t kor_smc_get_log(const kor_smc_t *s, kor_smc_log_t *log, int64_t cap);
/* trace of the LAST sweep: partner indices, stretch variate, lprob, proposal logprior, Xp, decision
 * decision: 0 dead, 1 prior -Inf, 2 failed prior-MH pre-test, 3 simulated+rejected, 4 accepted */
void kor_smc_get_trace(const kor_smc_t *s, int64_t *a, int64_t *b, double *z, double *lprob, double *lpi_p,
                       double *xp, uint8_t *decision, double *theta_p_soa);

/* ---- AIS: src/transition.jl, src/types.jl:51-75, src/KissABC.jl:35-80 ---- */
typedef struct kor_ais kor_ais_t;
int kor_ais_create(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model,
                   const kor_ais_config_t *cfg, int nthreads, kor_ais_t **out);
void kor_ais_destroy(kor_ais_t *s);
int kor_ais_init(kor_ais_t *s); /* src/KissABC.jl:50-61 */
/* one transition! of walker i; partners drawn from [cand_lo, cand_lo+cand_n) minus {i}.  returns 1 if accepted */
int kor_ais_transition(kor_ais_t *s, int64_t i, int64_t cand_lo, int64_t cand_n, uint32_t epoch);
/* one red/black sweep (two half-steps); epochs 2*sweep, 2*sweep+1 are consumed */
int kor_ais_sweep(kor_ais_t *s);
/* whole runs: out_samples is SoA d x nsamples */
int kor_ais_run_sequential(kor_ais_t *s, double *out_samples); /* reference schedule, KissABC.jl:66-80 */
void kor_ais_set_serial(kor_ais_t *s); /* as kor_smc_set_serial, for kor_ais_init + kor_ais_transition (call before kor_ais_init) */
int kor_ais_run_parallel(kor_ais_t *s, double *out_samples);   /* red/black schedule of the device path */
void kor_ais_get_state(const kor_ais_t *s, double *theta_soa, double *lp, double *ll);
void kor_ais_set_state(kor_ais_t *s, const double *theta_soa, const double *lp, const double *ll);
void kor_ais_get_counters(const kor_ais_t *s, int64_t *cost_evals, int64_t *accepted, int64_t *sweeps, int64_t *retries);
/* trace of the last sweep: move (1..3, 0 = none), partners, corr, proposal, new logdensity, e, decision
 * decision: 0 invalid-new (no variate consumed), 1 rejected, 2 accepted */
void kor_ais_get_trace(const kor_ais_t *s, uint8_t *move, int64_t *a, int64_t *b, int64_t *c, double *corr,
                       double *theta_p_soa, double *lp_p, double *ll_p, double *e, uint8_t *decision);

/* ---- ABCDE: src/smc.jl:352-428 (population Monte Carlo with differential-evolution moves, Jacobi update) ---- */
typedef struct {
    int64_t nparticles;    /* 50 */
    int64_t generations;   /* 20 */
    double eps_target;
    double alpha;          /* 0 <= alpha < 1 */
    double proposal_width; /* 1.0 */
    int32_t earlystop;     /* false */
    int32_t _pad;
} kor_abcde_config_t;
/* theta_out: d x N SoA of push_p'ed particles; cost_out: N; returns 0 or 1 (error text in kor_last_error) */
int kor_abcde_run(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model, const kor_abcde_config_t *cfg,
                  int nthreads, double *theta_out, double *cost_out, int32_t *reached, int64_t *nsim, int64_t *generations_done);

/* ---- pfilter: src/smc.jl:275-345 (quantile filter, bad particles redrawn from the good ones until they pass) ---- */
typedef struct {
    int64_t nparticles;
    double q;              /* 0.7 */
    double eff_tol;        /* 0.1 */
    double epstol;         /* -Inf */
    double proposal_width; /* 0.75 */
    int64_t max_iters;     /* 0 = Inf */
} kor_pfilter_config_t;
/* ref :276-279: N is raised to ceil((4d+1)/q) when N*q <= 4d */
int64_t kor_pfilter_nparticles(int64_t n, int d, double q);
int kor_pfilter_run(uint64_t seed, const kor_prior_t *prior, int d, const kor_model_t *model, const kor_pfilter_config_t *cfg,
                    int nthreads, double *theta_out, double *cost_out, double *eps_out, int64_t *iters, int64_t *nreps,
                    int64_t *cost_evals);

#ifdef __cplusplus
}
#endif
#endif
