// kabc_smc.cu -- smc(prior, cost; ...) on device.  Restates src/smc.jl:92-206 of KissABC.jl 3.0.1:
//   init                      :119-129   k_smc_init / k_smc_init_prior + k_smc_init_gk, k_smc_post_init
//   eps = quantile(Xs[alive]) :134       k_sel x2 (exact two-rank bucket select on FP64 keys + Statistics.jl type-7
//                                        interpolation); the first histogram of an iteration is accumulated by the
//                                        WRITERS of X during the previous sweep, so the usual iteration needs one pass
//   alive cut, flag, ESS      :135-142   k_cut
//   cyclic-tiling resample    :145-153   k_cut (decision + scan), k_compact (table of the surviving rows)
//   propose                   :160-167   tiles of k_smc_sweep_q, or k_smc_propose (phase A: own row through the resampling
//   prior-MH pre-test         :172-175    map, partners, proposal, prior tests; survivors go onto the work list)
//   simulate + accept         :176-189   chunks of k_smc_sweep_q, or k_smc_simulate_list / _lv / _gk (phase B: every warp full)
//   retry / stop rules        :156-159,192-198   post_sweep()/post_iter() run by the LAST block of the sweep kernel
// Every scalar that steers control flow lives in SmcCtrl in device memory; the host only reads `stop`.
//
// Layout.  The population is SHARDED: rank r of G owns the global particles [r P, (r+1) P), P = N/G, and holds only
// their state: th[k*P + li] (SoA FP64), X[li], lpi[li], alive[li].  Before every sweep the owner writes the rows a sweep
// may READ into its TABLE -- the surviving rows compacted in index order when the iteration resamples (the j-th alive
// particle of the population is entry j - off[r] of rank r's table, off = exclusive scan of the per-rank alive counts),
// all rows otherwise -- in the rank's peer arena (kabc_peer.cuh), one AoS row [theta | X | lpi] per particle (a partner
// costs one 16/32-byte read, the own row one 32/48-byte read).  A sweep reads rows only through their owner's table (peer
// loads for the other ranks): its own row idx[i] = idxalive[i mod n_alive] (the reference's cyclic tiling, :146-147) and the two
// partners a, b (:163-164); it writes only the state of its own shard.  So a rank receives O(P) rows per sweep whatever
// G is, the quantile / cut / scan work on the shard only, and the ranks meet in four flag barriers per iteration
// (histogram + sweep counters, candidate keys, alive counts, table complete) -- no NCCL on the data path, no host.
// Philox counters are keyed by the GLOBAL particle id: results are bit-identical for any G.
#include <ctime>
#include "kabc_host.hpp"
#include "kabc_gk.cuh"
#include "kabc_peer.cuh"

namespace kabc {

constexpr int SEL_LOG2_BINS = 12;
constexpr int SEL_BINS = 1 << SEL_LOG2_BINS;
constexpr int SEL_CAP = 4096; // candidates sorted exactly by one block
constexpr int SEL_THREADS = 512;
constexpr int SEL_INSTANCES = 2;   // k_sel launches per iteration (steady state needs one); whatever is left -- the third pass
                                   // of iteration 1, runs of ties -- is finished by the last block of the last instance alone
constexpr int SCAN_THREADS = 2048; // particles per block of the cut / compact kernels: 256 threads x 2 groups of 4; 2^20
                                   // particles are 512 blocks = ONE wave (1024 blocks were 1.15 waves: the 2nd wave cost as much)
constexpr int CUT_GROUPS = 2;
constexpr int CUT_THREADS = 256;

enum { SEL_HIST = 0, SEL_CAND = 1, SEL_KEYS = 2, SEL_FINAL = 3 };

// what rank s contributes to a cross-rank step, pushed into slot [barrier parity][s] of EVERY rank's arena
struct XSlot {
    unsigned int hist[SEL_BINS];
    unsigned long long cand[SEL_CAP];
    unsigned long long v[16];
};
enum { XV_ACC = 0, XV_WORK, XV_EVENTS, XV_MINKEY, XV_MAXKEY, XV_BELOW, XV_ABOVE, XV_COUNT, XV_ERR };

struct SmcCtrl {
    double eps, eps_prev, xmin, gamma;
    double win;                  // relative width of the window [eps (1-win), eps] the sweep histograms for the next quantile
    unsigned long long xmin_key; // min(Xs[alive]) as an ordered key, ref :136 -- exact, recomputed by every sweep
    // two-rank selection: the alive keys inside [klo,khi] are `cnt`, `below` alive keys are smaller, r0/r1 are the ranks
    unsigned long long klo, khi, v0key, v1key;
    long long below, cnt, r0, r1;
    unsigned long long h_klo, h_khi; // key range the histogram is accumulating (selection pass, or the running sweep)
    int h_shift, sel_state;
    long long n_alive; // number of alive particles (input of the next quantile)
    long long ess;     // ESS = sum(alive) right after the cut (what the reference prints)
    long long off[KABC_MAX_PEERS + 1]; // table entries [off[r], off[r+1]) live on rank r; off[G] = rows a sweep maps onto
    unsigned long long accepted, cost_evals, events;
    unsigned long long sw_accepted, sw_work, sw_events, sw_minkey, sw_maxkey, sw_below, sw_above; // this rank, this sweep
    long long iteration;
    unsigned int work_count, cand_count, epoch, lv_head, tile_head, tiles_done;
    unsigned int tk_sel, tk_cut, tk_compact, tk_sim, tk_misc;
    int flag, resample, stop, err, sweeps, retry_done, resampled_log;
    int honor_stop; // kabc_smc_run enqueues one iteration ahead: once `stop` is set the queued kernels do nothing
};

struct SmcParams { // launch constants
    long long N, P, lo; // population, shard size, first global index of the shard
    int d, TS;          // parameters, doubles per row of the table: [theta_0..theta_{d-1} | X | lpi] padded to an even count
    double alpha, mcmc_tol, epstol, r_epstol, min_r_ess, max_stretch, sqrt_np;
    long long mcmc_retrys;
    int max_iterations;
    int rank, world;
    unsigned int prop_cap; // queued sweep: most tiles in flight (claimed, not yet published) at any time; 0 = no limit
};

struct SmcTrace {
    long long *a, *b;
    double *z, *lprob, *lpip, *xp;
    unsigned char *dec;
};

struct SmcBufs {
    double *th, *X, *lpi; // state of the shard
    unsigned char *alive;
    // peer-visible block of every rank (arena + the handle's offset; the local buffer when G = 1):
    //   XSlot[2][G] | rows[P][TS] (one AoS row per particle: theta, X, lpi -- a partner costs one 16/32-byte read, the own
    //   row one 32/48-byte read) | tab_alive[P]
    unsigned char *xb[KABC_MAX_PEERS];
    long long o_th, o_alive; // byte offsets of the rows and of the alive flags inside xb[r]
    double *thp, *lpip;                  // proposals [d][P] (work-list path and trace)
    unsigned int *work, *blockcnt, *hist;
    unsigned int *fill;                  // entries of the work list written so far, per chunk of 256 (queued sweep)
    unsigned long long *cand;
    SmcCtrl *ctrl;
    kabc_smc_log_t *log;
    long long log_cap;
    SmcTrace tr;
    int trace_on;
};

__device__ __forceinline__ XSlot *xslot(const SmcBufs &B, const SmcParams &P, int r, int set, int src) {
    return reinterpret_cast<XSlot *>(B.xb[r]) + set * P.world + src;
}
__device__ __forceinline__ const double *tab_th(const SmcBufs &B, int r) { return reinterpret_cast<const double *>(B.xb[r] + B.o_th); }

// an iteration queued behind a stop (or after an error) must leave the state untouched
__device__ __forceinline__ bool smc_skip(const SmcCtrl *c) { return c->err || (c->honor_stop && c->stop); }

__device__ __forceinline__ unsigned long long dkey(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}
__device__ __forceinline__ int sel_shift(unsigned long long klo, unsigned long long khi) {
    const unsigned long long range = khi - klo;
    const int bl = range ? 64 - __clzll((long long)range) : 0;
    return bl > SEL_LOG2_BINS ? bl - SEL_LOG2_BINS : 0;
}
// bins a histogram over [klo,khi] with this shift can touch (the rest of the 4096 stay zero and are not exchanged)
__device__ __forceinline__ int sel_used_bins(unsigned long long klo, unsigned long long khi, int shift) {
    const unsigned long long nb = ((khi - klo) >> shift) + 1ull;
    return nb < (unsigned long long)SEL_BINS ? (int)nb : SEL_BINS;
}
__device__ __forceinline__ void raise_peer_error(SmcCtrl *c) {
    if (threadIdx.x == 0 && !c->err) c->err = KABC_ERR_PEER;
    __syncthreads();
}

// rank that holds table entry j (off is ascending, off[0] = 0) and the entry's index there
__device__ __forceinline__ int locate(const long long *off, int G, long long j, long long &jl) {
    int r = 0;
    for (int q = 1; q < G; ++q) r += (j >= off[q]) ? 1 : 0;
    jl = j - off[r];
    return r;
}

// ------------------------------------------------------------------ what the writers of X do for the next quantile
// Every alive particle's cost is final exactly once per sweep (in the propose phase if it does not reach the simulator,
// in the accept otherwise).  At that point it is binned for the NEXT iteration's quantile over the window
// [eps (1-win), eps] (every alive cost is < eps, ref :137), counted if it lies under / over the window, and folded into
// the exact minimum of the alive costs (ref :136).
struct FinalNote {
    unsigned long long kmin = ~0ull, kmax = 0ull;
    unsigned int below = 0, above = 0;
};
__device__ __forceinline__ void note_final(FinalNote &f, unsigned int *hist, unsigned long long klo, unsigned long long khi,
                                           int shift, double X) {
    const unsigned long long key = dkey(X);
    f.kmin = key < f.kmin ? key : f.kmin;
    f.kmax = key > f.kmax ? key : f.kmax;
    if (key < klo) f.below += 1;
    else if (key > khi) f.above += 1;
    else atomicAdd(&hist[(key - klo) >> shift], 1u);
}

// sweep constants + tallies of a sweep CTA in shared memory
struct SweepShared {
    long long off[KABC_MAX_PEERS + 1];
    double eps;
    unsigned long long hklo, hkhi, kmin;
    unsigned int acc, work, below, above, tile;
    int flag, hshift;
    uint32_t epoch;
};
__device__ __forceinline__ void note_final_shared(SweepShared &sh, unsigned int *hist, double X) {
    const unsigned long long key = dkey(X);
    // a 64-bit shared-memory min is a CAS loop (ATOMS.CAST.SPIN): only the rare keys under the running minimum try
    if (key < *reinterpret_cast<volatile unsigned long long *>(&sh.kmin)) atomicMin(&sh.kmin, key);
    if (key < sh.hklo) atomicAdd(&sh.below, 1u);
    else if (key > sh.hkhi) atomicAdd(&sh.above, 1u);
    else atomicAdd(&hist[(key - sh.hklo) >> sh.hshift], 1u);
}

// ------------------------------------------------------------------ init, ref src/smc.jl:119-129
template <int KIND, int PREC>
__global__ void __launch_bounds__(256)
k_smc_init(SmcBufs B, SmcParams P, DPriors pri, DModel m, RoundKeys rk) {
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long Pn = P.P;
    unsigned long long key = ~0ull, ev_u = 0;
    if (li < Pn) {
        const uint32_t id = (uint32_t)(P.lo + li);
        Stream st(rk, ST_PRIOR, id, 0u);
        bool ok = true;
#pragma unroll 1
        for (int k = 0; k < P.d; ++k) {
            double x;
            ok &= prior1_sample(pri.p[k], st, x);
            B.th[(long long)k * Pn + li] = x;
        }
        if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
        long long ev;
        double *thv = B.th;
        double X = cost_thread<KIND, PREC>(m, rk, ST_COST_INIT, id, 0u, [&](int k) { return thv[(long long)k * Pn + li]; }, ev);
        B.X[li] = X;
        B.lpi[li] = prior_logpdf_pushed(pri, [&](int k) { return thv[(long long)k * Pn + li]; });
        B.alive[li] = 1;
        key = dkey(X);
        ev_u = (unsigned long long)ev;
    }
    key = warp_min_u64(key);
    ev_u = warp_sum_u64(ev_u);
    if ((threadIdx.x & 31) == 0) {
        if (key != ~0ull) atomicMin(&B.ctrl->sw_minkey, key);
        if (ev_u) atomicAdd(&B.ctrl->sw_events, ev_u);
    }
}

__global__ void k_smc_init_prior(SmcBufs B, SmcParams P, DPriors pri, RoundKeys rk) {
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= P.P) return;
    const long long Pn = P.P;
    Stream st(rk, ST_PRIOR, (uint32_t)(P.lo + li), 0u);
    bool ok = true;
    for (int k = 0; k < P.d; ++k) {
        double x;
        ok &= prior1_sample(pri.p[k], st, x);
        B.th[(long long)k * Pn + li] = x;
    }
    if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
    double *thv = B.th;
    B.lpi[li] = prior_logpdf_pushed(pri, [&](int k) { return thv[(long long)k * Pn + li]; });
    B.alive[li] = 1;
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS, (PREC == KABC_F64 ? 1 : 3))
k_smc_init_gk(SmcBufs B, SmcParams P, DModel m, RoundKeys rk) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    const long long Pn = P.P;
    for (long long li = blockIdx.x; li < Pn; li += gridDim.x) {
        const double *th = B.th;
        double c = cost_gk_block<PREC>(m, rk, ST_COST_INIT, (uint32_t)(P.lo + li), 0u, pushk(m, 0, th[li]), pushk(m, 1, th[Pn + li]),
                                    pushk(m, 2, th[2 * Pn + li]), pushk(m, 3, th[3 * Pn + li]), gk_smem);
        if (threadIdx.x == 0) {
            B.X[li] = c;
            atomicMin(&B.ctrl->sw_minkey, dkey(c));
        }
    }
}

__global__ void k_smc_reset(SmcBufs B, SmcParams P) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int q = t; q < SEL_BINS; q += gridDim.x * blockDim.x) B.hist[q] = 0;
    if (t == 0) {
        SmcCtrl *c = B.ctrl;
        memset(c, 0, sizeof(SmcCtrl));
        c->eps = dinf(); c->eps_prev = dinf();
        c->xmin_key = ~0ull; c->sw_minkey = ~0ull;
        c->n_alive = P.N; c->ess = P.N;
        c->win = 0.5;
        c->cost_evals = (unsigned long long)P.N;
    }
}

// ranks r0, r1 of the two order statistics the type-7 quantile interpolates (Statistics.jl, ref :134) among n alive keys:
// aleph = n*p + (1-p); j = clamp(trunc(aleph),1,n-1); gamma = clamp(aleph-j,0,1).  Thread 0 of one block.
__device__ void sel_begin(SmcCtrl *c, const SmcParams &P, long long n) {
    if (n <= 0) { c->err = KABC_ERR_DEGENERATE; return; }
    const double aleph = xadd(xmul((double)n, P.alpha), xsub(1.0, P.alpha));
    long long j = (long long)aleph;
    if (j > n - 1) j = n - 1;
    if (j < 1) j = 1;
    double gamma = xsub(aleph, (double)j);
    gamma = gamma < 0.0 ? 0.0 : (gamma > 1.0 ? 1.0 : gamma);
    c->gamma = gamma;
    c->r0 = (n == 1) ? 0 : j - 1; // 0-based rank of v[j]
    c->r1 = (n == 1) ? 0 : j;     // 0-based rank of v[j+1]
}
// the next selection pass histograms every alive key (nothing is known about them)
__device__ void sel_restart_full(SmcCtrl *c) {
    c->klo = 0; c->khi = ~0ull; c->below = 0; c->cnt = c->n_alive;
    c->h_klo = 0; c->h_khi = ~0ull; c->h_shift = sel_shift(0, ~0ull);
    c->sel_state = SEL_HIST;
}

// after the init kernels (and after kabc_smc_set_state): fold the ranks' minimum / event count / alive count
__global__ void __launch_bounds__(32) k_smc_post_init(SmcBufs B, SmcParams P, XPeer x, int recount) {
    SmcCtrl *c = B.ctrl;
    const int set = (int)((*x.seq + 1ull) & 1ull);
    if (threadIdx.x < P.world) {
        XSlot *s = xslot(B, P, threadIdx.x, set, P.rank);
        s->v[XV_MINKEY] = c->sw_minkey; s->v[XV_EVENTS] = c->sw_events; s->v[XV_COUNT] = (unsigned long long)c->n_alive;
        s->v[XV_ERR] = (unsigned long long)c->err;
    }
    if (!xbarrier(x)) raise_peer_error(c);
    if (threadIdx.x == 0) {
        unsigned long long mk = ~0ull, ev = 0, n = 0;
        for (int r = 0; r < P.world; ++r) {
            const XSlot *s = xslot(B, P, P.rank, set, r);
            mk = s->v[XV_MINKEY] < mk ? s->v[XV_MINKEY] : mk;
            ev += s->v[XV_EVENTS]; n += s->v[XV_COUNT];
            if (!c->err && s->v[XV_ERR]) c->err = (int)s->v[XV_ERR]; // a prior that cannot be sampled on one rank stops all
        }
        c->xmin_key = mk;
        if (recount) { c->n_alive = (long long)n; c->ess = (long long)n; }
        else { c->events = ev; c->n_alive = P.N; c->ess = P.N; }
        c->sw_minkey = ~0ull; c->sw_events = 0; c->sw_accepted = 0;
        sel_begin(c, P, c->n_alive);
        sel_restart_full(c);
    }
}

// ------------------------------------------------------------------ quantile, ref src/smc.jl:134 (Statistics type 7)
// Exact selection of the two adjacent order statistics v[j], v[j+1] by range narrowing: histogram the alive keys
// inside [klo,khi] into 4096 equal-width key bins (every rank its shard; the per-rank histograms are pushed to every
// peer and summed redundantly), keep the bins holding the two ranks, repeat; once at most SEL_CAP keys remain they are
// compacted, pushed to every rank and sorted by one block.  Keys are the order-preserving u64 image of the doubles, so
// bins, ranks and the result are exact (no floating point in the selection).
struct SelRange {
    unsigned long long klo, khi;
    long long below, cnt, r0, r1;
    int shift;
};
// block-wide (NT threads): loc[] = this thread's SEL_BINS/NT consecutive bin counts; narrows R to the bins holding r0, r1
template <int NT>
__device__ void sel_scan_narrow(const unsigned int (&loc)[SEL_BINS / NT], SelRange &R, unsigned int *s_scan, unsigned long long *s_res) {
    constexpr int PER = SEL_BINS / NT;
    unsigned int sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) sum += loc[q];
    s_scan[threadIdx.x] = sum;
    if (threadIdx.x < 4) s_res[threadIdx.x] = 0;
    __syncthreads();
    for (int o = 1; o < NT; o <<= 1) {
        unsigned int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0u;
        __syncthreads();
        s_scan[threadIdx.x] += v;
        __syncthreads();
    }
    const long long excl = (long long)s_scan[threadIdx.x] - sum;
    const long long t0 = R.r0 - R.below, t1 = R.r1 - R.below;
    long long cum = excl;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const long long nxt = cum + (long long)loc[q];
        if (t0 >= cum && t0 < nxt) { s_res[0] = threadIdx.x * PER + q; s_res[1] = (unsigned long long)cum; }
        if (t1 >= cum && t1 < nxt) { s_res[2] = threadIdx.x * PER + q; s_res[3] = (unsigned long long)nxt; }
        cum = nxt;
    }
    __syncthreads();
    const unsigned long long b0 = s_res[0], before = s_res[1], b1 = s_res[2], through = s_res[3];
    const unsigned long long mask = R.shift ? ((1ull << R.shift) - 1ull) : 0ull;
    const unsigned long long off_end = (b1 << R.shift) | mask;
    const unsigned long long new_khi = off_end > R.khi - R.klo ? R.khi : R.klo + off_end;
    R.klo = R.klo + (b0 << R.shift);
    R.khi = new_khi;
    R.below += (long long)before;
    R.cnt = (long long)(through - before);
    __syncthreads();
}

// Executed by ONE block per rank (NT threads) once the shard's histogram over [h_klo,h_khi] is complete in B.hist:
// exchange it (+ the scalars in v[], prepared by the caller in the own XSlot of every rank), sum, narrow, decide what the
// next selection step is.  from_sweep: the histogram was accumulated by the sweep over the predicted window; the
// numbers of alive keys under / over the window come with it and the prediction may have failed.
// gmin/gmax (pass mode): extreme in-range keys over all ranks -- tighten the range (this is what ends runs of ties).
template <int NT>
__device__ void sel_after_hist(SmcBufs &B, const SmcParams &P, const XPeer &x, int set, bool from_sweep,
                               unsigned long long below_sw, unsigned long long above_sw, unsigned long long gmin,
                               unsigned long long gmax, unsigned int *s_scan, unsigned long long *s_res) {
    constexpr int PER = SEL_BINS / NT;
    SmcCtrl *c = B.ctrl;
    unsigned int loc[PER];
    const int nb_used = sel_used_bins(c->h_klo, c->h_khi, c->h_shift);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int bin = threadIdx.x * PER + q;
        unsigned int s = 0;
        if (bin < nb_used)
            for (int r = 0; r < P.world; ++r) s += xslot(B, P, P.rank, set, r)->hist[bin];
        loc[q] = s;
    }
    SelRange R;
    R.klo = c->h_klo; R.khi = c->h_khi; R.shift = c->h_shift; R.r0 = c->r0; R.r1 = c->r1;
    R.below = from_sweep ? (long long)below_sw : c->below;
    R.cnt = from_sweep ? c->n_alive - (long long)below_sw - (long long)above_sw : c->cnt;
    const bool failed = from_sweep && (above_sw > 0 || R.r0 < (long long)below_sw);
    __syncthreads();
    if (failed) { // a rank lies outside the predicted window: histogram [true min, top] in the next pass
        R.klo = c->xmin_key; R.khi = above_sw > 0 ? ~0ull : c->h_khi; R.below = 0; R.cnt = c->n_alive; R.shift = 1;
        if (R.khi < R.klo) { R.klo = 0; R.khi = ~0ull; }
    } else {
        sel_scan_narrow<NT>(loc, R, s_scan, s_res);
        if (!from_sweep) {
            if (gmin > R.klo && gmin <= R.khi) R.klo = gmin;
            if (gmax < R.khi && gmax >= R.klo) R.khi = gmax;
        }
    }
    if (threadIdx.x == 0) {
        c->klo = R.klo; c->khi = R.khi; c->below = R.below; c->cnt = R.cnt;
        if (!failed && (R.shift == 0 || R.klo == R.khi)) { // bins were single keys (or one key is left): these ARE v[j], v[j+1]
            c->v0key = R.klo; c->v1key = R.khi;
            c->sel_state = SEL_KEYS;
        } else if (!failed && R.cnt <= SEL_CAP) {
            c->sel_state = SEL_CAND;
        } else {
            c->h_klo = R.klo; c->h_khi = R.khi; c->h_shift = sel_shift(R.klo, R.khi);
            c->sel_state = SEL_HIST;
        }
    }
    __syncthreads();
}

// closes the selection and opens the iteration, ref :132-141.  One thread.
__device__ void sel_finalize(SmcCtrl *c) {
    const double a = dunkey(c->v0key), b = dunkey(c->v1key), g = c->gamma;
    double eps;
    if (dfinite(a) && dfinite(b)) eps = xadd(a, xmul(g, xsub(b, a)));
    else eps = xadd(xmul(xsub(1.0, g), a), xmul(g, b));
    c->iteration += 1;
    c->eps_prev = c->eps;
    c->eps = eps;
    c->xmin = dunkey(c->xmin_key);
    c->flag = (eps > c->xmin) ? 0 : 1; // ref :136-141
    c->sweeps = 0; c->retry_done = 0; c->accepted = 0; c->resampled_log = 0;
    // window of the NEXT quantile: four times the last relative step of eps, at most the top binade
    double win = 0.5;
    if (dfinite(c->eps_prev) && c->eps_prev > 0.0 && eps > 0.0 && eps < c->eps_prev) {
        win = xmul(4.0, xsub(1.0, xdiv(eps, c->eps_prev)));
        win = win < 0.015625 ? 0.015625 : (win > 0.5 ? 0.5 : win);
    }
    c->win = win;
    c->sel_state = SEL_FINAL;
}

// 4 consecutive particles per thread: one uchar4 + two double2 loads
__device__ __forceinline__ void load4(const unsigned char *alive, const double *X, long long base, long long N,
                                      unsigned int (&a)[4], double (&x)[4]) {
    if (base + 3 < N) {
        const uchar4 av = *reinterpret_cast<const uchar4 *>(alive + base);
        const double2 x0 = *reinterpret_cast<const double2 *>(X + base), x1 = *reinterpret_cast<const double2 *>(X + base + 2);
        a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
        x[0] = x0.x; x[1] = x0.y; x[2] = x1.x; x[3] = x1.y;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            a[q] = (base + q < N) ? alive[base + q] : 0u;
            x[q] = (base + q < N) ? X[base + q] : 0.0;
        }
    }
}

// one histogram pass of block `blk` of `nblk` over the shard: alive keys inside [h_klo,h_khi] -> B.hist, extreme keys
__device__ void sel_pass_hist(SmcBufs &B, const SmcParams &P, unsigned int *sh, int blk, int nblk) {
    SmcCtrl *c = B.ctrl;
    const unsigned long long klo = c->h_klo, khi = c->h_khi;
    const int shift = c->h_shift;
    for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) sh[q] = 0;
    __syncthreads();
    unsigned long long kmin = ~0ull, kmax = 0ull;
    const long long stride = (long long)nblk * blockDim.x * 4;
    for (long long base = ((long long)blk * blockDim.x + threadIdx.x) * 4; base < P.P; base += stride) {
        unsigned int a[4];
        double xv[4];
        load4(B.alive, B.X, base, P.P, a, xv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!a[q]) continue;
            const unsigned long long key = dkey(xv[q]);
            if (key >= klo && key <= khi) {
                atomicAdd(&sh[(key - klo) >> shift], 1u);
                kmin = key < kmin ? key : kmin;
                kmax = key > kmax ? key : kmax;
            }
        }
    }
    kmin = warp_min_u64(kmin);
    kmax = warp_max_u64(kmax);
    if ((threadIdx.x & 31) == 0) {
        if (kmin != ~0ull) atomicMin(&c->sw_minkey, kmin);
        if (kmax != 0ull) atomicMax(&c->sw_maxkey, kmax);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) {
        const unsigned int v = sh[q];
        if (v) atomicAdd(&B.hist[q], v);
    }
}
// one compaction pass: the alive keys inside [klo,khi] (at most SEL_CAP over all ranks) -> B.cand
__device__ void sel_pass_cand(SmcBufs &B, const SmcParams &P, int blk, int nblk) {
    SmcCtrl *c = B.ctrl;
    const unsigned long long klo = c->klo, khi = c->khi;
    const long long stride = (long long)nblk * blockDim.x * 4;
    const long long nloop = (P.P + stride - 1) / stride;
    for (long long it = 0; it < nloop; ++it) {
        const long long base = it * stride + ((long long)blk * blockDim.x + threadIdx.x) * 4;
        unsigned int a[4] = {0u, 0u, 0u, 0u};
        double xv[4] = {0.0, 0.0, 0.0, 0.0};
        if (base < P.P) load4(B.alive, B.X, base, P.P, a, xv);
        unsigned long long keys[4];
        bool ins[4], any = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            keys[q] = dkey(xv[q]);
            ins[q] = a[q] && keys[q] >= klo && keys[q] <= khi;
            any |= ins[q];
        }
        if (__ballot_sync(0xffffffffu, any) == 0) continue; // candidates are rare (<= 4096 of N)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool in = ins[q];
            const unsigned int ball = __ballot_sync(0xffffffffu, in);
            if (ball) {
                const unsigned int lane = threadIdx.x & 31;
                unsigned int pos = 0;
                if (lane == 0) pos = atomicAdd(&c->cand_count, (unsigned int)__popc(ball));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                const unsigned int at = pos + __popc(ball & ((1u << lane) - 1u));
                if (in && at < SEL_CAP) B.cand[at] = keys[q];
            }
        }
    }
}

// the last block of a selection step: publish the shard's contribution, meet the other ranks, consume.
// s_buf: SEL_CAP u64 of shared memory.
__device__ void sel_publish_consume(SmcBufs &B, const SmcParams &P, const XPeer &x, int st, unsigned long long *s_buf,
                                    unsigned int *s_scan, unsigned long long *s_res) {
    SmcCtrl *c = B.ctrl;
    const int set = (int)((*x.seq + 1ull) & 1ull);
    if (st == SEL_HIST) {
        const int nb_used = sel_used_bins(c->h_klo, c->h_khi, c->h_shift);
        for (int r = 0; r < P.world; ++r) {
            XSlot *s = xslot(B, P, r, set, P.rank);
            for (int q = threadIdx.x; q < nb_used; q += blockDim.x) s->hist[q] = __ldcg(&B.hist[q]);
            if (threadIdx.x == 0) { s->v[XV_MINKEY] = c->sw_minkey; s->v[XV_MAXKEY] = c->sw_maxkey; }
        }
        __syncthreads();
        for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) B.hist[q] = 0;
        if (!xbarrier(x)) raise_peer_error(c);
        unsigned long long gmin = ~0ull, gmax = 0ull;
        for (int r = 0; r < P.world; ++r) {
            const XSlot *s = xslot(B, P, P.rank, set, r);
            gmin = s->v[XV_MINKEY] < gmin ? s->v[XV_MINKEY] : gmin;
            gmax = s->v[XV_MAXKEY] > gmax ? s->v[XV_MAXKEY] : gmax;
        }
        if (threadIdx.x == 0) { c->sw_minkey = ~0ull; c->sw_maxkey = 0ull; }
        sel_after_hist<SEL_THREADS>(B, P, x, set, false, 0, 0, gmin, gmax, s_scan, s_res);
    } else { // SEL_CAND
        const unsigned int n_loc = c->cand_count < (unsigned)SEL_CAP ? c->cand_count : (unsigned)SEL_CAP;
        for (int r = 0; r < P.world; ++r) {
            XSlot *s = xslot(B, P, r, set, P.rank);
            for (unsigned int q = threadIdx.x; q < n_loc; q += blockDim.x) s->cand[q] = __ldcg(&B.cand[q]);
            if (threadIdx.x == 0) s->v[XV_COUNT] = n_loc;
        }
        if (!xbarrier(x)) raise_peer_error(c);
        // concatenate the ranks' lists, sort, pick
        int n = 0;
        for (int r = 0; r < P.world; ++r) {
            const XSlot *s = xslot(B, P, P.rank, set, r);
            const unsigned long long nr64 = s->v[XV_COUNT];
            const int nr = nr64 > (unsigned long long)SEL_CAP ? SEL_CAP : (int)nr64;
            for (int q = threadIdx.x; q < nr; q += blockDim.x)
                if (n + q < SEL_CAP) s_buf[n + q] = s->cand[q];
            n += nr;
        }
        if (n != (int)c->cnt || n > SEL_CAP) { // cannot happen: the histogram counted exactly these keys
            if (threadIdx.x == 0 && !c->err) c->err = KABC_ERR_STATE;
            n = n > SEL_CAP ? SEL_CAP : n;
        }
        int npad = 2;
        while (npad < n) npad <<= 1;
        __syncthreads();
        for (int q = n + threadIdx.x; q < npad; q += blockDim.x) s_buf[q] = ~0ull;
        __syncthreads();
        for (int k = 2; k <= npad; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const bool up = (i & k) == 0;
                    const unsigned long long a = s_buf[i], b = s_buf[l];
                    if ((a > b) == up) { s_buf[i] = b; s_buf[l] = a; }
                }
                __syncthreads();
            }
        if (threadIdx.x == 0) {
            long long i0 = c->r0 - c->below, i1 = c->r1 - c->below;
            if (i0 < 0 || i0 >= n || i1 < 0 || i1 >= n) { if (!c->err) c->err = KABC_ERR_STATE; i0 = 0; i1 = 0; }
            c->v0key = s_buf[i0];
            c->v1key = s_buf[i1];
            c->cand_count = 0;
            c->sel_state = SEL_KEYS;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SEL_THREADS) k_sel(SmcBufs B, SmcParams P, XPeer x, int last_instance) {
    __shared__ unsigned long long s_buf[SEL_CAP]; // histogram (u32 view) or candidate keys
    __shared__ unsigned int s_scan[SEL_THREADS];
    __shared__ unsigned long long s_res[4];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    const int st = c->sel_state;
    if (st == SEL_FINAL) return;
    if (st == SEL_KEYS) { // decided at the end of the last sweep (one key left): only the bookkeeping remains
        if (blockIdx.x == 0 && threadIdx.x == 0) sel_finalize(c);
        return;
    }
    if (st == SEL_HIST) sel_pass_hist(B, P, reinterpret_cast<unsigned int *>(s_buf), blockIdx.x, gridDim.x);
    else sel_pass_cand(B, P, blockIdx.x, gridDim.x);
    if (!last_block(&c->tk_sel)) return;
    sel_publish_consume(B, P, x, st, s_buf, s_scan, s_res);
    if (last_instance) { // massive ties / pathological spread: this block alone keeps narrowing over its shard
        for (;;) {
            const int st2 = c->sel_state;
            if (st2 == SEL_KEYS || c->err) break;
            if (st2 == SEL_HIST) sel_pass_hist(B, P, reinterpret_cast<unsigned int *>(s_buf), 0, 1);
            else sel_pass_cand(B, P, 0, 1);
            __threadfence();
            __syncthreads();
            sel_publish_consume(B, P, x, st2, s_buf, s_scan, s_res);
        }
    }
    if (threadIdx.x == 0) {
        c->tk_sel = 0;
        if (c->sel_state == SEL_KEYS && !c->err) sel_finalize(c);
    }
}

// ------------------------------------------------------------------ alive cut + ESS + resample decision, ref :136-147
// block b owns local particles [2048 b, 2048 b + 2048): 256 threads x 8 consecutive particles
__device__ __forceinline__ unsigned int block_excl_scan_256(unsigned int v, unsigned int *s_w, unsigned int &total) {
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned int w = threadIdx.x < (CUT_THREADS / 32) ? s_w[threadIdx.x] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= o) wi += t;
        }
        s_w[threadIdx.x] = wi - w;          // exclusive warp offsets
        if (threadIdx.x == 31) s_w[32] = wi; // block total
    }
    __syncthreads();
    total = s_w[32];
    const unsigned int r = incl - v + s_w[warp];
    __syncthreads();
    return r;
}

// what a sweep needs before it starts; thread 0 of the block that closed the previous step.  `resample`: the sweep maps
// onto the compacted tables (offsets = scan of the ranks' alive counts in cnt[]), else onto all rows in place.
__device__ void sweep_setup(SmcCtrl *c, const SmcParams &P, int resample, const unsigned long long *cnt) {
    long long run = 0;
    for (int r = 0; r < P.world; ++r) {
        c->off[r] = run;
        run += resample ? (long long)cnt[r] : P.P;
    }
    c->off[P.world] = run;
    c->resample = resample;
    c->sw_accepted = 0; c->sw_work = 0; c->sw_events = 0; c->sw_minkey = ~0ull; c->sw_maxkey = 0ull; c->sw_below = 0; c->sw_above = 0;
    c->work_count = 0; c->lv_head = 0; c->tile_head = 0; c->tiles_done = 0;
    // window of the next quantile: every alive cost is < eps (<= with the flag), ref :137-139
    const double eps = c->eps;
    const unsigned long long khi = dkey(eps);
    unsigned long long klo = 0;
    if (dfinite(eps) && eps > 0.0) {
        klo = dkey(xmul(eps, xsub(1.0, c->win)));
        if (klo > khi) klo = 0;
    }
    c->h_klo = klo; c->h_khi = khi; c->h_shift = sel_shift(klo, khi);
}

__global__ void __launch_bounds__(CUT_THREADS) k_cut(SmcBufs B, SmcParams P, XPeer x, int nblocks) {
    __shared__ unsigned int s_w[33];
    __shared__ unsigned long long s_cnt[KABC_MAX_PEERS];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    if (c->sel_state != SEL_FINAL) { // the selection did not finish (cannot happen: k_sel loops until it does)
        if (blockIdx.x == 0 && threadIdx.x == 0 && !c->err) c->err = KABC_ERR_STATE;
        return;
    }
    const double *X = B.X;
    const double eps = c->eps;
    const int flag = c->flag;
    unsigned int cnt = 0;
#pragma unroll
    for (int g = 0; g < CUT_GROUPS; ++g) { // thread t owns particles [8t, 8t+8) of the block's 2048
        const long long base = ((long long)blockIdx.x * CUT_THREADS + threadIdx.x) * (4 * CUT_GROUPS) + 4 * g;
        if (base + 3 < P.P) {
            const double2 x0 = *reinterpret_cast<const double2 *>(X + base), x1 = *reinterpret_cast<const double2 *>(X + base + 2);
            uchar4 a;
            a.x = flag ? (x0.x <= eps) : (x0.x < eps); a.y = flag ? (x0.y <= eps) : (x0.y < eps);
            a.z = flag ? (x1.x <= eps) : (x1.x < eps); a.w = flag ? (x1.y <= eps) : (x1.y < eps);
            *reinterpret_cast<uchar4 *>(B.alive + base) = a;
            cnt += a.x + a.y + a.z + a.w;
        } else {
            for (int q = 0; q < 4; ++q)
                if (base + q < P.P) {
                    const double xv = X[base + q];
                    const unsigned int a = flag ? (xv <= eps) : (xv < eps);
                    B.alive[base + q] = (unsigned char)a;
                    cnt += a;
                }
        }
    }
    unsigned int total;
    block_excl_scan_256(cnt, s_w, total);
    if (threadIdx.x == 0) B.blockcnt[blockIdx.x] = total;
    if (!last_block(&c->tk_cut)) return;
    // last block: exclusive scan of the per-block counts (in place); total = alive particles of the shard
    unsigned long long carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += CUT_THREADS) {
        const int q = b0 + threadIdx.x;
        const unsigned int v = q < nblocks ? __ldcg(&B.blockcnt[q]) : 0u;
        unsigned int chunk;
        const unsigned int excl = block_excl_scan_256(v, s_w, chunk);
        if (q < nblocks) B.blockcnt[q] = (unsigned int)(carry + excl);
        carry += chunk;
    }
    // all-gather of the G counts: global scan offsets + ESS
    const int set = (int)((*x.seq + 1ull) & 1ull);
    if (threadIdx.x < P.world) xslot(B, P, threadIdx.x, set, P.rank)->v[XV_COUNT] = carry;
    if (!xbarrier(x)) raise_peer_error(c);
    if (threadIdx.x < P.world) s_cnt[threadIdx.x] = xslot(B, P, P.rank, set, threadIdx.x)->v[XV_COUNT];
    __syncthreads();
    if (threadIdx.x == 0) {
        long long ess = 0;
        for (int r = 0; r < P.world; ++r) ess += (long long)s_cnt[r];
        c->ess = ess;
        c->tk_cut = 0;
        // ref :145  alpha*ESS <= nparticles*min_r_ess, FP64, exactly these operands
        const int resample = xmul(P.alpha, (double)ess) <= xmul((double)P.N, P.min_r_ess);
        if (resample && ess == 0) c->err = KABC_ERR_DEGENERATE;
        c->n_alive = resample ? P.N : ess; // ref :151-152: after resampling everything is alive
        if (resample) c->resampled_log = 1;
        sweep_setup(c, P, resample, s_cnt);
    }
}

// ------------------------------------------------------------------ the table a sweep reads, ref :146-152
// resampling: entry (blockcnt[b] + rank inside the block) of the table <- the alive rows of the shard in index order
// (idxalive = (1:N)[alive] restricted to the shard), then `alive .= true`; otherwise entry li <- row li.
__global__ void __launch_bounds__(CUT_THREADS, 4) k_compact(SmcBufs B, SmcParams P, XPeer x, int force_identity, int barrier) {
    __shared__ unsigned int s_w[33];
    SmcCtrl *c = B.ctrl;
    if (force_identity ? (c->err == KABC_ERR_PEER) : (smc_skip(c) || c->retry_done)) return; // get_state runs after `stop` too
    const int resample = force_identity ? 0 : c->resample;
    const long long Pn = P.P;
    constexpr int PT = 4 * CUT_GROUPS; // particles per thread
    const long long base = ((long long)blockIdx.x * CUT_THREADS + threadIdx.x) * PT;
    { // the work list of the coming sweep starts empty
        const long long nfill = (Pn + 255) / 256 + 1, gt = (long long)blockIdx.x * CUT_THREADS + threadIdx.x;
        for (long long q = gt; q < nfill; q += (long long)gridDim.x * CUT_THREADS) B.fill[q] = 0;
    }
    double *t_th = reinterpret_cast<double *>(B.xb[P.rank] + B.o_th);
    unsigned char *t_alive = B.xb[P.rank] + B.o_alive;
    unsigned int a[PT];
    const bool fullv = base + PT - 1 < Pn;
    if (fullv) {
#pragma unroll
        for (int g = 0; g < CUT_GROUPS; ++g) {
            const uchar4 av = *reinterpret_cast<const uchar4 *>(B.alive + base + 4 * g);
            a[4 * g] = av.x; a[4 * g + 1] = av.y; a[4 * g + 2] = av.z; a[4 * g + 3] = av.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < PT; ++q) a[q] = (base + q < Pn) ? B.alive[base + q] : 0u;
    }
    unsigned int pos = 0;
    if (resample) {
        unsigned int total, mine = 0;
#pragma unroll
        for (int q = 0; q < PT; ++q) mine += a[q];
        pos = B.blockcnt[blockIdx.x] + block_excl_scan_256(mine, s_w, total);
    }
    if (fullv && P.d == 2 && (Pn & 1) == 0) { // the common shape: vector loads of the SoA state, two 128-bit stores per 32-byte row
        double t0[PT], t1[PT], xx[PT], ll[PT];
#pragma unroll
        for (int g = 0; g < PT / 2; ++g) {
            const double2 v0 = *reinterpret_cast<const double2 *>(B.th + base + 2 * g), v1 = *reinterpret_cast<const double2 *>(B.th + Pn + base + 2 * g);
            const double2 v2 = *reinterpret_cast<const double2 *>(B.X + base + 2 * g), v3 = *reinterpret_cast<const double2 *>(B.lpi + base + 2 * g);
            t0[2 * g] = v0.x; t0[2 * g + 1] = v0.y; t1[2 * g] = v1.x; t1[2 * g + 1] = v1.y;
            xx[2 * g] = v2.x; xx[2 * g + 1] = v2.y; ll[2 * g] = v3.x; ll[2 * g + 1] = v3.y;
        }
#pragma unroll
        for (int q = 0; q < PT; ++q) {
            if (resample && !a[q]) continue;
            const long long e = resample ? (long long)pos++ : base + q;
            double2 *row = reinterpret_cast<double2 *>(t_th + e * 4);
            row[0] = make_double2(t0[q], t1[q]);
            row[1] = make_double2(xx[q], ll[q]);
            if (!resample) t_alive[e] = (unsigned char)a[q];
        }
    } else {
#pragma unroll 1
        for (int q = 0; q < PT; ++q) {
            const long long li = base + q;
            if (li >= Pn) continue;
            if (resample && !a[q]) continue;
            const long long e = resample ? (long long)pos++ : li;
            double *row = t_th + e * P.TS;
            for (int k = 0; k < P.d; ++k) row[k] = B.th[(long long)k * Pn + li];
            row[P.d] = B.X[li];
            row[P.d + 1] = B.lpi[li];
            if (!resample) t_alive[e] = (unsigned char)a[q];
        }
    }
    if (resample) {
        if (fullv) {
#pragma unroll
            for (int g = 0; g < CUT_GROUPS; ++g) *reinterpret_cast<uchar4 *>(B.alive + base + 4 * g) = make_uchar4(1, 1, 1, 1);
        } else {
            for (int q = 0; q < PT; ++q)
                if (base + q < Pn) B.alive[base + q] = 1;
        }
    }
    if (!barrier || P.world == 1) return;
    if (!last_block(&c->tk_compact)) return;
    if (!xbarrier(x)) raise_peer_error(c); // every table is complete before any rank's sweep reads it
    if (threadIdx.x == 0) c->tk_compact = 0;
}

// ------------------------------------------------------------------ sweep / iteration bookkeeping
// ref :192 (`accepted >= mcmc_tol*nparticles && break`): folds the sweep's counters of every rank into the control block
__device__ void post_sweep(SmcCtrl *c, const SmcParams &P, unsigned long long acc, unsigned long long work, unsigned long long ev,
                           unsigned long long mk) {
    c->accepted += acc;
    c->cost_evals += work;
    c->events += ev;
    c->xmin_key = mk; // every alive cost was seen by this sweep
    c->sweeps += 1;
    c->epoch += 1;
    if ((double)c->accepted >= xmul(P.mcmc_tol, (double)P.N)) c->retry_done = 1;
}
// closes an iteration, ref :194-198
__device__ void post_iter(SmcBufs &B, const SmcParams &P) {
    SmcCtrl *c = B.ctrl;
    if (c->err) { c->stop = -1; return; }
    const double eps = c->eps, epsv = c->eps_prev;
    int stop = 0;
    if (xmul(2.0, fabs(xsub(epsv, eps))) < xmul(P.r_epstol, xadd(fabs(epsv), fabs(eps)))) stop = 1;
    else if (eps <= P.epstol) stop = 2;
    else if ((double)c->accepted < xmul(P.mcmc_tol, (double)P.N)) stop = 3;
    else if (P.max_iterations > 0 && c->iteration >= P.max_iterations) stop = 4;
    c->stop = stop;
    const long long it = c->iteration;
    if (it >= 1 && it <= B.log_cap) {
        kabc_smc_log_t &L = B.log[it - 1];
        L.iteration = it; L.eps = eps; L.n_alive = c->ess; L.flag = c->flag; L.resampled = c->resampled_log;
        L.accepted = (long long)c->accepted; L.cost_evals = (long long)c->cost_evals; L.sweeps = c->sweeps;
    }
}

// What the last block (NT threads) of a sweep kernel does: exchange the sweep's counters and the histogram of the final
// costs with the other ranks, close the sweep (and, close_iter, the iteration), and start the next iteration's quantile.
template <int NT>
__device__ void sweep_finish(SmcBufs &B, const SmcParams &P, const XPeer &x, int close_iter, unsigned int *s_scan,
                             unsigned long long *s_res) {
    SmcCtrl *c = B.ctrl;
    const int set = (int)((*x.seq + 1ull) & 1ull);
    const int nb_used = sel_used_bins(c->h_klo, c->h_khi, c->h_shift);
    for (int r = 0; r < P.world; ++r) {
        XSlot *s = xslot(B, P, r, set, P.rank);
        for (int q = threadIdx.x; q < nb_used; q += NT) s->hist[q] = __ldcg(&B.hist[q]);
        if (threadIdx.x == 0) {
            s->v[XV_ACC] = c->sw_accepted; s->v[XV_WORK] = c->sw_work; s->v[XV_EVENTS] = c->sw_events;
            s->v[XV_MINKEY] = c->sw_minkey; s->v[XV_MAXKEY] = c->sw_maxkey; s->v[XV_BELOW] = c->sw_below; s->v[XV_ABOVE] = c->sw_above;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < SEL_BINS; q += NT) B.hist[q] = 0;
    if (!xbarrier(x)) raise_peer_error(c);
    unsigned long long acc = 0, work = 0, ev = 0, mk = ~0ull, below = 0, above = 0;
    for (int r = 0; r < P.world; ++r) {
        const XSlot *s = xslot(B, P, P.rank, set, r);
        acc += s->v[XV_ACC]; work += s->v[XV_WORK]; ev += s->v[XV_EVENTS];
        mk = s->v[XV_MINKEY] < mk ? s->v[XV_MINKEY] : mk;
        below += s->v[XV_BELOW]; above += s->v[XV_ABOVE];
    }
    if (threadIdx.x == 0) {
        c->tk_sim = 0;
        post_sweep(c, P, acc, work, ev, mk);
        if (close_iter) post_iter(B, P);
        // the next step reads the state as it is now: a retry sweep of the same iteration through the identity map (ref
        // :159: no second resampling), the next iteration through its quantile
        sel_begin(c, P, c->n_alive);
    }
    __syncthreads();
    // the histogram was taken over the window [h_klo,h_khi] sweep_setup chose; consume it BEFORE the window is re-armed
    if (!c->err) sel_after_hist<NT>(B, P, x, set, true, below, above, 0, ~0ull, s_scan, s_res);
    if (threadIdx.x == 0) {
        unsigned long long none[KABC_MAX_PEERS] = {};
        const int st = c->sel_state;
        const unsigned long long hk0 = c->h_klo, hk1 = c->h_khi;
        const int hs = c->h_shift;
        sweep_setup(c, P, 0, none); // identity map + fresh counters for a possible retry sweep (same eps, same window)
        if (st == SEL_HIST) { c->h_klo = hk0; c->h_khi = hk1; c->h_shift = hs; } // ... unless a selection pass is pending
    }
}
__global__ void k_post_iter(SmcBufs B, SmcParams P) { post_iter(B, P); }
__global__ void k_set_honor_stop(SmcBufs B, int v) { B.ctrl->honor_stop = v; }

// ------------------------------------------------------------------ propose, ref :147-152 (own row), :160-167, :172-175
// One particle: its own row after the (possible) resampling -- table entry i mod n through the owners' tables -- goes
// into the state of the shard; if it is alive, partners a, b (through the same map), stretch variate, proposal, prior,
// prior-MH pre-test.  Returns true when the proposal goes on to the simulator.
template <int DM>
struct Proposed {
    double thp[DM], lpip, Xi;
};
// theta of table row t (first d doubles of the row); own = true also returns X and lpi (doubles d and d+1)
template <int DM>
__device__ __forceinline__ void load_row(const double *t, int d, double (&row)[DM], bool own, double &X, double &lp) {
    if (DM == 2 && d == 2) {
        const double2 v = reinterpret_cast<const double2 *>(t)[0];
        row[0] = v.x; row[1] = v.y;
        if (own) { const double2 w = reinterpret_cast<const double2 *>(t)[1]; X = w.x; lp = w.y; }
    } else if (DM <= 4 && d >= 3) { // d = 3, 4: rows of 6 doubles
        const double2 v0 = reinterpret_cast<const double2 *>(t)[0], v1 = reinterpret_cast<const double2 *>(t)[1];
        row[0] = v0.x; row[1] = v0.y; row[2 % DM] = v1.x;
        if (DM == 4) row[3 % DM] = v1.y;
        if (own) {
            const double2 v2 = reinterpret_cast<const double2 *>(t)[2];
            if (d == 3) { X = v1.y; lp = v2.x; } else { X = v2.x; lp = v2.y; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < DM; ++k) row[k] = k < d ? t[k] : 0.0;
        if (own) { X = t[d]; lp = t[d + 1]; }
    }
}
template <int DM>
__device__ __forceinline__ bool smc_propose_one(const SmcBufs &B, const SmcParams &P, const SmcCtrl *c, const DPriors &pri,
                                                const RoundKeys &rk, const long long *off, long long li, Proposed<DM> &out,
                                                bool &alive_out) {
    const long long Pn = P.P, i = P.lo + li;
    const int G = P.world;
    const int resample = c->resample;
    const long long n_src = off[G];
    const uint32_t epoch = c->epoch;
    int dec = 0;
    long long a = -1, b = -1;
    double z = dnan(), lprob = dnan(), lpip = dnan();
    bool pass = false;
    // the particle's own row after the (possible) resampling
    long long el;
    const int ro = locate(off, G, resample ? (long long)((unsigned int)i % (unsigned int)n_src) : i, el);
    double row[DM], Xi = 0.0, lpi_i = 0.0, unused0, unused1;
    load_row<DM>(tab_th(B, ro) + el * P.TS, P.d, row, true, Xi, lpi_i);
    const bool alive_i = resample ? true : (B.alive[li] != 0);
    if (resample) {
#pragma unroll
        for (int k = 0; k < DM; ++k)
            if (k < P.d) B.th[(long long)k * Pn + li] = row[k];
        B.X[li] = Xi;
        B.lpi[li] = lpi_i;
    }
    out.Xi = Xi;
    alive_out = alive_i;
    if (alive_i) {
        Stream st(rk, ST_PROPOSE, (uint32_t)i, epoch);
        a = i; b = i;
        while (a == i) a = (long long)index_of(st.next(), (uint32_t)P.N);
        while (b == i || b == a) b = (long long)index_of(st.next(), (uint32_t)P.N);
        long long ea, eb;
        const int ra = locate(off, G, resample ? (long long)((unsigned int)a % (unsigned int)n_src) : a, ea);
        const int rb = locate(off, G, resample ? (long long)((unsigned int)b % (unsigned int)n_src) : b, eb);
        double pa[DM], pb[DM];
        load_row<DM>(tab_th(B, ra) + ea * P.TS, P.d, pa, false, unused0, unused1);
        load_row<DM>(tab_th(B, rb) + eb * P.TS, P.d, pb, false, unused0, unused1);
        z = next_normal(st);
        const double sc = xdiv(xmul(P.max_stretch, z), P.sqrt_np);
#pragma unroll
        for (int k = 0; k < DM; ++k) out.thp[k] = k < P.d ? xadd(row[k], xmul(xsub(pb[k], pa[k]), sc)) : 0.0;
        const uint32_t wu = st.next();
        lpip = prior_logpdf_pushed(pri, [&](int k) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < DM; ++q) v = (q == k) ? out.thp[q] : v;
            return v;
        });
        if (lpip < 0.0 && !dfinite(lpip)) dec = 1;
        else {
            // ref :174-175  lM = min(lpip - lpi + logcorr, 0); proceed iff log(rand) < lM.  Julia's min propagates NaN and
            // `lprob < NaN` is false, so a NaN ratio skips the proposal.  log(u) < 0 always (u < 1), so the logarithm is
            // only evaluated when lM < 0 (or when tracing).
            const double dl = xadd(xsub(lpip, lpi_i), 0.0);
            if (dl != dl) dec = 2;
            else {
                const double lM = fmin(dl, 0.0);
                pass = true;
                if (!(lM >= 0.0) || B.trace_on) {
                    lprob = xlog(u01(wu));
                    pass = lprob < lM;
                }
                if (!pass) dec = 2;
            }
        }
        if (B.trace_on && lprob != lprob) lprob = xlog(u01(wu));
        out.lpip = lpip;
    }
    if (B.trace_on) {
        B.tr.a[i] = a; B.tr.b[i] = b; B.tr.z[i] = z; B.tr.lprob[i] = lprob; B.tr.lpip[i] = lpip;
        B.tr.dec[i] = (unsigned char)dec; B.tr.xp[i] = dnan();
        for (int k = 0; k < P.d; ++k) {
            double v = dnan();
#pragma unroll
            for (int q = 0; q < DM; ++q) v = (alive_i && q == k) ? out.thp[q] : v;
            B.thp[(long long)k * Pn + li] = v;
        }
    }
    return pass;
}

// ------------------------------------------------------------------ accept, ref :176-189
template <typename F>
__device__ __forceinline__ bool smc_accept(SmcBufs &B, const SmcParams &P, double eps, int flag, long long li, double Xp, double lpip,
                                           F thp_of) {
    const bool reject = flag ? (Xp > eps) : (Xp >= eps);
    if (!reject) {
        for (int k = 0; k < P.d; ++k) B.th[(long long)k * P.P + li] = thp_of(k);
        B.X[li] = Xp;
        B.lpi[li] = lpip;
    }
    if (B.trace_on) { B.tr.xp[P.lo + li] = Xp; B.tr.dec[P.lo + li] = reject ? 3 : 4; }
    return !reject;
}

// per-thread tallies of a sweep kernel -> one set of global atomics per CTA
struct SweepTally {
    unsigned int acc = 0, work = 0;
    unsigned long long events = 0;
    FinalNote f;
};
__device__ __forceinline__ void tally_flush(SmcCtrl *c, SweepTally &t) {
    __shared__ unsigned int s_t[4];
    __shared__ unsigned long long s_u[3];
    if (threadIdx.x == 0) { s_t[0] = 0; s_t[1] = 0; s_t[2] = 0; s_t[3] = 0; s_u[0] = 0; s_u[1] = ~0ull; s_u[2] = 0ull; }
    __syncthreads();
    const unsigned int acc = (unsigned int)warp_sum_u64(t.acc), work = (unsigned int)warp_sum_u64(t.work);
    const unsigned int below = (unsigned int)warp_sum_u64(t.f.below), above = (unsigned int)warp_sum_u64(t.f.above);
    const unsigned long long ev = warp_sum_u64(t.events), kmin = warp_min_u64(t.f.kmin), kmax = warp_max_u64(t.f.kmax);
    if ((threadIdx.x & 31) == 0) {
        if (acc) atomicAdd(&s_t[0], acc);
        if (work) atomicAdd(&s_t[1], work);
        if (below) atomicAdd(&s_t[2], below);
        if (above) atomicAdd(&s_t[3], above);
        if (ev) atomicAdd(&s_u[0], ev);
        atomicMin(&s_u[1], kmin);
        atomicMax(&s_u[2], kmax);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_t[0]) atomicAdd(&c->sw_accepted, (unsigned long long)s_t[0]);
        if (s_t[1]) atomicAdd(&c->sw_work, (unsigned long long)s_t[1]);
        if (s_t[2]) atomicAdd(&c->sw_below, (unsigned long long)s_t[2]);
        if (s_t[3]) atomicAdd(&c->sw_above, (unsigned long long)s_t[3]);
        if (s_u[0]) atomicAdd(&c->sw_events, s_u[0]);
        if (s_u[1] != ~0ull) atomicMin(&c->sw_minkey, s_u[1]);
        if (s_u[2] != 0ull) atomicMax(&c->sw_maxkey, s_u[2]);
    }
}

// ------------------------------------------------------------------ propose kernel: phase A of a sweep, ref :160-167 (+ :172-175)
template <int DM>
__global__ void __launch_bounds__(256, 5)
k_smc_propose(SmcBufs B, SmcParams P, DPriors pri, RoundKeys rk) {
    __shared__ long long s_off[KABC_MAX_PEERS + 1];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || c->retry_done) return;
    if (threadIdx.x <= P.world) s_off[threadIdx.x] = c->off[threadIdx.x];
    __syncthreads();
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    SweepTally tl;
    bool pass = false, alive_i = false;
    if (li < P.P) {
        Proposed<DM> pr;
        pass = smc_propose_one<DM>(B, P, c, pri, rk, s_off, li, pr, alive_i);
        if (alive_i && !pass) note_final(tl.f, B.hist, c->h_klo, c->h_khi, c->h_shift, pr.Xi);
        if (pass) {
#pragma unroll
            for (int k = 0; k < DM; ++k)
                if (k < P.d) B.thp[(long long)k * P.P + li] = pr.thp[k];
            B.lpip[li] = pr.lpip;
        }
    }
    // block-aggregated append to the work list: ONE global atomic per CTA
    __shared__ unsigned int s_cnt[8], s_base;
    const unsigned int ball = __ballot_sync(0xffffffffu, pass);
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_cnt[warp] = __popc(ball);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const unsigned int v = s_cnt[w]; s_cnt[w] = tot; tot += v; }
        s_base = tot ? atomicAdd(&c->work_count, tot) : 0u;
    }
    __syncthreads();
    if (pass) B.work[s_base + s_cnt[warp] + __popc(ball & ((1u << lane) - 1u))] = (unsigned int)li;
    tally_flush(c, tl);
}

// Phase B of a sweep for the thread-per-particle simulators (ref :168-191): persistent CTAs claim chunks of 256 work items
// from an atomic head, one particle per thread, every warp full (the list holds only proposals that passed the prior tests).
// Tried and rejected (numbers in DESIGN.md section 5): L lanes of a warp sharing one particle's draws (finer chunks, but
// +6 / +9 / +16 % time at L = 2 / 4 / 8) and one fused propose+simulate kernel with a per-CTA survivor queue (+8 %).
template <int KIND, int PREC>
__global__ void __launch_bounds__(256, 6) k_smc_simulate_list(SmcBufs B, SmcParams P, XPeer x, DModel m, RoundKeys rk, int close_iter) {
    __shared__ unsigned int s_scan[256];
    __shared__ unsigned long long s_res[4];
    __shared__ SweepShared sh; // sweep constants and tallies live in shared memory: the register budget is the draw loop's
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || c->retry_done) return;
    constexpr unsigned int PER = 256;
    if (threadIdx.x == 0) {
        sh.eps = c->eps; sh.flag = c->flag; sh.epoch = c->epoch;
        sh.hklo = c->h_klo; sh.hkhi = c->h_khi; sh.hshift = c->h_shift;
        sh.kmin = ~0ull; sh.acc = 0; sh.work = 0; sh.below = 0; sh.above = 0;
    }
    const unsigned int nwork = c->work_count;
    unsigned long long events = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sh.tile = atomicAdd(&c->lv_head, 1u);
        __syncthreads();
        const unsigned int base = sh.tile * PER;
        if (base >= nwork) break;
        const unsigned int w = base + threadIdx.x;
        const bool valid = w < nwork;
        const long long li = valid ? (long long)B.work[w] : 0;
        const long long Pn = P.P;
        const double *thp = B.thp;
        long long ev = 0;
        if ((w & ~31u) >= nwork) continue; // whole warp past the end of the list
        const double Xp = valid ? cost_thread<KIND, PREC>(m, rk, ST_COST, (uint32_t)(P.lo + li), sh.epoch, [&](int k) { return thp[(long long)k * Pn + li]; }, ev) : 0.0;
        if (valid) {
            const double Xold = B.X[li];
            const bool ok = smc_accept(B, P, sh.eps, sh.flag, li, Xp, B.lpip[li], [&](int k) { return thp[(long long)k * Pn + li]; });
            note_final_shared(sh, B.hist, ok ? Xp : Xold);
            atomicAdd(&sh.work, 1u);
            if (ok) atomicAdd(&sh.acc, 1u);
            if (KIND == KABC_MODEL_LV_SSA) events += (unsigned long long)ev;
        }
    }
    __syncthreads();
    if (KIND == KABC_MODEL_LV_SSA) {
        events = warp_sum_u64(events);
        if ((threadIdx.x & 31) == 0 && events) atomicAdd(&c->sw_events, events);
    }
    if (threadIdx.x == 0) {
        if (sh.acc) atomicAdd(&c->sw_accepted, (unsigned long long)sh.acc);
        if (sh.work) atomicAdd(&c->sw_work, (unsigned long long)sh.work);
        if (sh.below) atomicAdd(&c->sw_below, (unsigned long long)sh.below);
        if (sh.above) atomicAdd(&c->sw_above, (unsigned long long)sh.above);
        if (sh.kmin != ~0ull) atomicMin(&c->sw_minkey, sh.kmin);
    }
    if (last_block(&c->tk_sim)) sweep_finish<256>(B, P, x, close_iter, s_scan, s_res);
}

// ------------------------------------------------------------------ the queued sweep (thread-per-particle simulators)
// ONE persistent kernel runs both phases of a sweep.  Phase A units are TILES of 256 consecutive particles (propose: own
// row and partner gathers -- NVLink peer loads on a multi-GPU job --, FP64 proposal arithmetic, prior tests); their
// survivors are appended to a global work list.  Phase B units are CHUNKS of 256 consecutive list entries (simulate +
// accept, every warp full).  A CTA that finds a complete chunk simulates it; otherwise it proposes the next tile; so
// proposals are produced just ahead of their consumption, spread over the whole kernel, and their (peer) load latency is
// covered by the other CTAs' draw loops, while the balance between CTAs is that of a dynamically claimed chunk.
// Phase A of EVERY particle still reads only the frozen table and phase B only writes the state of its own particle, so
// the result is the Jacobi update of the reference (:160-191) whatever the interleaving.
// Protocol: a tile's entries are written, fenced, then counted into fill[chunk] (atomicAdd) and tiles_done; chunk h may be
// claimed (CAS on the head) once fill[h] == 256, or == the remainder when all tiles are done.
enum { QACT_SIM = 1, QACT_PROP = 2, QACT_EXIT = 3, QACT_RETRY = 4 };
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int KIND, int PREC, int DM>
__global__ void __launch_bounds__(256, 6)
k_smc_sweep_q(SmcBufs B, SmcParams P, XPeer x, DPriors pri, DModel m, RoundKeys rk, int close_iter) {
    __shared__ unsigned int s_scan[256];
    __shared__ unsigned long long s_res[4];
    __shared__ SweepShared sh;
    __shared__ unsigned int s_cnt[8], s_base, s_act, s_unit, s_len;
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || c->retry_done) return;
    if (threadIdx.x <= P.world) sh.off[threadIdx.x] = c->off[threadIdx.x];
    if (threadIdx.x == 0) {
        sh.eps = c->eps; sh.flag = c->flag; sh.epoch = c->epoch;
        sh.hklo = c->h_klo; sh.hkhi = c->h_khi; sh.hshift = c->h_shift;
        sh.kmin = ~0ull; sh.acc = 0; sh.work = 0; sh.below = 0; sh.above = 0;
    }
    const long long Pn = P.P;
    const unsigned int ntiles = (unsigned int)((Pn + 255) / 256);
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long t_wait0 = 0;
    unsigned int my_chunk = 0xFFFFFFFFu; // thread 0: the chunk this CTA has claimed and not simulated yet
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) {
            // Claims are plain atomicAdds (no retry loop: 888 CTAs claim thousands of units per sweep).  A CTA claims the next
            // chunk first; while that chunk is incomplete it produces: it proposes tiles, whose survivors fill the chunks in
            // order, its own included.  Claims past the end of the list are recognised once every tile is published.
            unsigned int act = QACT_RETRY, unit = 0, len = 0, claimed = 0;
            if (my_chunk == 0xFFFFFFFFu) my_chunk = atomicAdd(&c->lv_head, 1u);
            const unsigned int h = my_chunk;
            const unsigned int done = ld_acquire_gpu_u32(&c->tiles_done);
            const unsigned int total = ld_acquire_gpu_u32(&c->work_count);
            const unsigned int f = ld_acquire_gpu_u32(&B.fill[h < ntiles ? h : ntiles]);
            const bool all_done = done == ntiles;
            const unsigned int rem = (all_done && total > h * 256u) ? (total - h * 256u) : 0u;
            if (h < ntiles && (f == 256u || (all_done && rem > 0u && rem < 256u && f == rem))) {
                act = QACT_SIM; unit = h; len = f;
                my_chunk = 0xFFFFFFFFu;
            } else if (all_done && total <= h * 256u) {
                act = QACT_EXIT;
            } else if ((claimed = ld_acquire_gpu_u32(&c->tile_head)) < ntiles && !(P.prop_cap && claimed - done >= P.prop_cap)) {
                const unsigned int t = atomicAdd(&c->tile_head, 1u);
                if (t < ntiles) { act = QACT_PROP; unit = t; }
            } else if (claimed < ntiles) {
                // multi-GPU: a proposal is three peer loads per particle; when every CTA proposes at once (the start of a
                // sweep) the link queues them all and no chunk completes until all do.  Bounding the tiles in flight lets the
                // first chunks complete -- and their simulation start -- while the rest of the proposals stream behind them.
                __nanosleep(200);
            } else { // every tile is claimed, some are still in flight on other CTAs: bounded wait
                const unsigned long long now = global_timer_ns();
                if (t_wait0 == 0) t_wait0 = now;
                if (now - t_wait0 > 5000000000ull) { if (!c->err) c->err = KABC_ERR_STATE; act = QACT_EXIT; }
                __nanosleep(100);
            }
            if (act != QACT_RETRY) t_wait0 = 0;
            s_act = act; s_unit = unit; s_len = len;
        }
        __syncthreads();
        const unsigned int act = s_act, unit = s_unit, len = s_len;
        __syncthreads();
        if (act == QACT_EXIT) break;
        if (act == QACT_PROP) {
            const long long li = (long long)unit * 256 + threadIdx.x;
            bool pass = false, alive_i = false;
            if (li < Pn) {
                Proposed<DM> pr;
                pass = smc_propose_one<DM>(B, P, c, pri, rk, sh.off, li, pr, alive_i);
                if (alive_i && !pass) note_final_shared(sh, B.hist, pr.Xi);
                if (pass) {
#pragma unroll
                    for (int k = 0; k < DM; ++k)
                        if (k < P.d) B.thp[(long long)k * Pn + li] = pr.thp[k];
                    B.lpip[li] = pr.lpip;
                }
            }
            const unsigned int ball = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) s_cnt[warp] = __popc(ball);
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned int tot = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { const unsigned int v = s_cnt[w]; s_cnt[w] = tot; tot += v; }
                s_base = tot ? atomicAdd(&c->work_count, tot) : 0u;
                s_len = tot;
            }
            __syncthreads();
            const unsigned int base = s_base, tot = s_len;
            if (pass) B.work[base + s_cnt[warp] + __popc(ball & ((1u << lane) - 1u))] = (unsigned int)li;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) { // publish: per-chunk fill counts, then the tile
                if (tot) {
                    const unsigned int c0 = base / 256u, c1 = (base + tot - 1u) / 256u;
                    if (c0 == c1) atomicAdd(&B.fill[c0], tot);
                    else {
                        const unsigned int first = (c0 + 1u) * 256u - base;
                        atomicAdd(&B.fill[c0], first);
                        atomicAdd(&B.fill[c1], tot - first);
                    }
                }
                __threadfence();
                atomicAdd(&c->tiles_done, 1u);
            }
        } else if (act == QACT_SIM) {
            if (threadIdx.x < len && (threadIdx.x & ~31u) < len) {
                const long long li = (long long)__ldcg(&B.work[unit * 256u + threadIdx.x]);
                const double *thp = B.thp;
                long long ev = 0;
                const double Xp = cost_thread<KIND, PREC>(m, rk, ST_COST, (uint32_t)(P.lo + li), sh.epoch,
                                                          [&](int k) { return __ldcg(&thp[(long long)k * Pn + li]); }, ev);
                const double Xold = __ldcg(&B.X[li]);
                const bool ok = smc_accept(B, P, sh.eps, sh.flag, li, Xp, __ldcg(&B.lpip[li]), [&](int k) { return __ldcg(&thp[(long long)k * Pn + li]); });
                note_final_shared(sh, B.hist, ok ? Xp : Xold);
                if (ok) atomicAdd(&sh.acc, 1u);
            }
            if (threadIdx.x == 0) sh.work += len;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh.acc) atomicAdd(&c->sw_accepted, (unsigned long long)sh.acc);
        if (sh.work) atomicAdd(&c->sw_work, (unsigned long long)sh.work);
        if (sh.below) atomicAdd(&c->sw_below, (unsigned long long)sh.below);
        if (sh.above) atomicAdd(&c->sw_above, (unsigned long long)sh.above);
        if (sh.kmin != ~0ull) atomicMin(&c->sw_minkey, sh.kmin);
    }
    if (last_block(&c->tk_sim)) sweep_finish<256>(B, P, x, close_iter, s_scan, s_res);
}

// Lotka-Volterra sweep: persistent lanes.  Event counts per trajectory differ by orders of magnitude, so a lane whose
// trajectory ended immediately pulls the next work item (warp-aggregated atomic on the work-list head) instead of
// idling until the slowest lane of its warp finishes.  Per-particle arithmetic is unchanged (LvSim), so results do
// not depend on the schedule.
constexpr int LV_CHUNK = 32; // events between refill checks
template <int PREC>
__global__ void __launch_bounds__(256) k_smc_simulate_lv(SmcBufs B, SmcParams P, XPeer x, DModel m, RoundKeys rk, int close_iter) {
    __shared__ unsigned int s_scan[256];
    __shared__ unsigned long long s_res[4];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || c->retry_done) return;
    const unsigned int nwork = c->work_count;
    const long long Pn = P.P;
    const uint32_t epoch = c->epoch;
    const double eps = c->eps;
    const int flag = c->flag;
    const unsigned long long hklo = c->h_klo, hkhi = c->h_khi;
    const int hshift = c->h_shift;
    const unsigned int lane = threadIdx.x & 31;
    LvSim<PREC != KABC_F64> sim;
    long long li = -1;
    bool have = false, exhausted = false;
    SweepTally tl;
    for (;;) {
        const unsigned int need = __ballot_sync(0xffffffffu, !have && !exhausted);
        if (need) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&c->lv_head, (unsigned int)__popc(need));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!have && !exhausted) {
                const unsigned int w = base + __popc(need & ((1u << lane) - 1u));
                if (w < nwork) {
                    li = B.work[w];
                    sim.init(m, ST_COST, (uint32_t)(P.lo + li), epoch, pushk(m, 0, B.thp[li]), pushk(m, 1, B.thp[Pn + li]),
                             pushk(m, 2, B.thp[2 * Pn + li]));
                    have = true;
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__ballot_sync(0xffffffffu, have)) break;
        bool fin = false;
        for (int e = 0; e < LV_CHUNK; ++e) {
            if (have && !fin) fin = sim.step(m, rk);
            if (!__ballot_sync(0xffffffffu, have && !fin)) break;
        }
        if (have && fin) {
            const double Xp = sim.result(), Xold = B.X[li];
            const double *thp = B.thp;
            const bool ok = smc_accept(B, P, eps, flag, li, Xp, B.lpip[li], [&](int k) { return thp[(long long)k * Pn + li]; });
            tl.work += 1;
            tl.acc += ok ? 1u : 0u;
            tl.events += (unsigned long long)sim.ev;
            note_final(tl.f, B.hist, hklo, hkhi, hshift, ok ? Xp : Xold);
            have = false;
        }
    }
    tally_flush(c, tl);
    if (last_block(&c->tk_sim)) sweep_finish<256>(B, P, x, close_iter, s_scan, s_res);
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS, (PREC == KABC_F64 ? 1 : 3)) k_smc_simulate_gk(SmcBufs B, SmcParams P, XPeer x, DModel m, RoundKeys rk, int close_iter) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    __shared__ unsigned int s_scan[GK_THREADS];
    __shared__ unsigned long long s_res[4];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || c->retry_done) return;
    const unsigned int nwork = c->work_count;
    const long long Pn = P.P;
    const double eps = c->eps;
    const int flag = c->flag;
    SweepTally tl;
    for (unsigned int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const long long li = B.work[w];
        const double *thp = B.thp;
        double Xp = cost_gk_block<PREC>(m, rk, ST_COST, (uint32_t)(P.lo + li), c->epoch, pushk(m, 0, thp[li]), pushk(m, 1, thp[Pn + li]),
                                     pushk(m, 2, thp[2 * Pn + li]), pushk(m, 3, thp[3 * Pn + li]), gk_smem);
        if (threadIdx.x == 0) {
            const double Xold = B.X[li];
            const bool ok = smc_accept(B, P, eps, flag, li, Xp, B.lpip[li], [&](int k) { return thp[(long long)k * Pn + li]; });
            tl.work += 1;
            tl.acc += ok ? 1u : 0u;
            note_final(tl.f, B.hist, c->h_klo, c->h_khi, c->h_shift, ok ? Xp : Xold);
        }
    }
    tally_flush(c, tl);
    if (last_block(&c->tk_sim)) sweep_finish<GK_THREADS>(B, P, x, close_iter, s_scan, s_res);
}

// recount after kabc_smc_set_state: alive count and minimum of the shard (k_smc_post_init folds the ranks)
__global__ void k_recount(SmcBufs B, SmcParams P) {
    __shared__ unsigned long long s_n, s_min;
    if (threadIdx.x == 0) { s_n = 0; s_min = ~0ull; }
    __syncthreads();
    unsigned long long n = 0, mk = ~0ull;
    for (long long i = threadIdx.x; i < P.P; i += blockDim.x)
        if (B.alive[i]) {
            n += 1;
            const unsigned long long k = dkey(B.X[i]);
            mk = k < mk ? k : mk;
        }
    n = warp_sum_u64(n);
    mk = warp_min_u64(mk);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_n, n); atomicMin(&s_min, mk); }
    __syncthreads();
    if (threadIdx.x == 0) {
        B.ctrl->n_alive = (long long)s_n;
        B.ctrl->sw_minkey = s_min;
        B.ctrl->sw_events = 0;
    }
    for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) B.hist[q] = 0;
}

// the whole population (rows of every rank, read from the identity tables) in the caller's layout: theta[k*N + i], ...
__global__ void __launch_bounds__(256) k_gather_full(SmcBufs B, SmcParams P, XPeer x, double *th, double *X, double *lpi,
                                                     unsigned char *alive) {
    const long long N = (th || X || lpi || alive) ? P.N : 0; // a rank that wants nothing only takes part in the barrier
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / P.P);
        const long long e = i - (long long)r * P.P;
        const double *t = tab_th(B, r) + e * P.TS;
        if (th) for (int k = 0; k < P.d; ++k) th[(long long)k * N + i] = t[k];
        if (X) X[i] = t[P.d];
        if (lpi) lpi[i] = t[P.d + 1];
        if (alive) alive[i] = (B.xb[r] + B.o_alive)[e];
    }
    if (P.world == 1) return;
    if (!last_block(&B.ctrl->tk_misc)) return;
    if (!xbarrier(x)) raise_peer_error(B.ctrl); // nobody rebuilds a table another rank is still reading
    if (threadIdx.x == 0) B.ctrl->tk_misc = 0;
}

} // namespace kabc

using namespace kabc;

// =================================================================== host side
struct kabc_smc {
    kabc_ctx *ctx = nullptr;
    DPriors pri;
    DModel model;
    SmcParams P;
    kabc_smc_config_t cfg;
    SmcBufs B;
    XPeer X;
    DevBuf<double> state, thp, lpip; // state = [th (d*P) | X (P) | lpi (P)]
    DevBuf<unsigned char> alive, xlocal; // xlocal: the peer-visible block when there are no peers (G = 1)
    bool in_arena = false;
    DevBuf<unsigned int> work, blockcnt, hist, fill;
    DevBuf<unsigned long long> cand;
    DevBuf<SmcCtrl> ctrl;
    DevBuf<kabc_smc_log_t> log;
    // trace
    DevBuf<long long> ta, tb;
    DevBuf<double> tz, tlprob, tlpip, txp;
    DevBuf<unsigned char> tdec;
    SmcCtrl *h_ctrl = nullptr; // pinned
    bool inited = false;
    long long launches = 0;
    int nblocks_scan = 0;
    // one whole iteration (selection + cut + table + sweep, all control flow on the device) captured as a CUDA graph: a
    // single launch instead of six, so the short kernels run back to back even when the host is not ahead of the device
    cudaGraphExec_t iter_graph = nullptr;
    int graph_kernels = 0;
    bool graph_ok = true;
    // optional warm per-kernel timing of one iteration (kabc_smc_profile_iteration)
    std::vector<cudaEvent_t> *prof = nullptr;
    void mark() {
        if (!prof) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, ctx->stream);
        prof->push_back(e);
    }
};

static int smc_check_cfg(const kabc_smc_config_t *cfg, int d) {
    // ref src/smc.jl:107-118, same order
    if (!(cfg->min_r_ess > 0)) return set_error(KABC_ERR_INVALID_ARG, "min_r_ess must be > 0.");
    if (!(cfg->mcmc_retrys >= 0)) return set_error(KABC_ERR_INVALID_ARG, "mcmc_retrys must be >= 0.");
    if (!(cfg->alpha > 0)) return set_error(KABC_ERR_INVALID_ARG, "alpha must be > 0.");
    if (!(cfg->r_epstol >= 0)) return set_error(KABC_ERR_INVALID_ARG, "r_epstol must be >= 0");
    if (!(cfg->mcmc_tol >= 0)) return set_error(KABC_ERR_INVALID_ARG, "mcmc_tol must be >= 0");
    if (!(cfg->max_stretch > 1)) return set_error(KABC_ERR_INVALID_ARG, "max_stretch must be > 1");
    double mn = cfg->alpha < cfg->min_r_ess ? cfg->alpha : cfg->min_r_ess;
    long long min_np = (long long)ceil(3.0 * (double)d / mn);
    if (cfg->nparticles < min_np) return set_error(KABC_ERR_INVALID_ARG, "nparticles must be >= %lld.", min_np);
    if (cfg->nparticles > 0xFFFFFFFFll) return set_error(KABC_ERR_INVALID_ARG, "nparticles must be < 2^32");
    return KABC_OK;
}

#define SMC_LAUNCHED(s, n) do { (s)->launches += (n); (s)->ctx->launches += (n); } while (0)

static inline int table_stride(int d) { return (d + 2 + 1) & ~1; } // [theta | X | lpi], even: rows stay 16-byte aligned
static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
// layout of the peer-visible block of a handle (identical on every rank)
struct XLayout { size_t o_th, o_alive, bytes; };
static XLayout smc_xlayout(long long P, int d, int world) {
    XLayout L;
    size_t o = align256(sizeof(XSlot) * 2 * (size_t)world);
    L.o_th = o; o += align256((size_t)P * table_stride(d) * 8);
    L.o_alive = o; o += align256((size_t)P);
    L.bytes = o;
    return L;
}

template <int KIND>
static void smc_launch_list_t(kabc_smc *s, int ci) {
    static const bool full_grid = [] { const char *e = getenv("KABC_SIM_GRID"); return e && e[0] == 'f'; }(); // A/B: one chunk per CTA
    const long long need = (s->P.P + 255) / 256, cap = (long long)s->ctx->sm_count * 6;
    const unsigned blocks = (unsigned)((need < cap || full_grid) ? need : cap);
    cudaStream_t st = s->ctx->stream;
    if (s->model.precision == KABC_F64)
        k_smc_simulate_list<KIND, KABC_F64><<<blocks, 256, 0, st>>>(s->B, s->P, s->X, s->model, s->ctx->rk, ci);
    else
        k_smc_simulate_list<KIND, KABC_F32_ACC64><<<blocks, 256, 0, st>>>(s->B, s->P, s->X, s->model, s->ctx->rk, ci);
}

template <int KIND>
static void smc_launch_queued_t(kabc_smc *s, int ci) {
    const long long need = (s->P.P + 255) / 256, cap = (long long)s->ctx->sm_count * 6;
    const unsigned blocks = (unsigned)(need < cap ? need : cap);
    cudaStream_t st = s->ctx->stream;
    const bool f64 = s->model.precision == KABC_F64;
    if constexpr (KIND == KABC_MODEL_DETERMINISTIC) { // the only registered simulator with a free dimension
        if (s->P.d > 2) {
            k_smc_sweep_q<KIND, KABC_F64, KABC_MAX_DIM><<<blocks, 256, 0, st>>>(s->B, s->P, s->X, s->pri, s->model, s->ctx->rk, ci);
            return;
        }
    }
    if (f64) k_smc_sweep_q<KIND, KABC_F64, 2><<<blocks, 256, 0, st>>>(s->B, s->P, s->X, s->pri, s->model, s->ctx->rk, ci);
    else k_smc_sweep_q<KIND, KABC_F32_ACC64, 2><<<blocks, 256, 0, st>>>(s->B, s->P, s->X, s->pri, s->model, s->ctx->rk, ci);
}

template <int KIND>
static void smc_launch_init_t(kabc_smc *s) {
    const unsigned blocks = (unsigned)((s->P.P + 255) / 256);
    if (s->model.precision == KABC_F64)
        k_smc_init<KIND, KABC_F64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->pri, s->model, s->ctx->rk);
    else
        k_smc_init<KIND, KABC_F32_ACC64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->pri, s->model, s->ctx->rk);
    SMC_LAUNCHED(s, 1);
}

static int smc_gk_grid(kabc_smc *s, size_t &smem) {
    smem = gk_smem_bytes(s->model.n_draws, s->model.precision);
    long long cap = (long long)s->ctx->sm_count * gk_blocks_per_sm(s->model.n_draws, s->model.precision);
    long long n = s->P.P;
    return (int)(n < cap ? n : cap);
}

static int smc_enqueue_init(kabc_smc *s) {
    NvtxRange nv("kabc:smc:init");
    kabc_ctx *ctx = s->ctx;
    k_smc_reset<<<4, 1024, 0, ctx->stream>>>(s->B, s->P);
    SMC_LAUNCHED(s, 1);
    switch (s->model.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: smc_launch_init_t<KABC_MODEL_NORMAL_MEANSTD>(s); break;
    case KABC_MODEL_MA2_AUTOCOV: smc_launch_init_t<KABC_MODEL_MA2_AUTOCOV>(s); break;
    case KABC_MODEL_LV_SSA: smc_launch_init_t<KABC_MODEL_LV_SSA>(s); break;
    case KABC_MODEL_DETERMINISTIC: smc_launch_init_t<KABC_MODEL_DETERMINISTIC>(s); break;
    case KABC_MODEL_SOCKS: smc_launch_init_t<KABC_MODEL_SOCKS>(s); break;
    case KABC_MODEL_GK_OCTILE: {
        k_smc_init_prior<<<(unsigned)((s->P.P + 255) / 256), 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk);
        size_t smem;
        int grid = smc_gk_grid(s, smem);
        if (s->model.precision == KABC_F64)
            k_smc_init_gk<KABC_F64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->model, ctx->rk);
        else
            k_smc_init_gk<KABC_F32_ACC64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->model, ctx->rk);
        SMC_LAUNCHED(s, 2);
        break;
    }
    }
    KABC_CUDA_TRY(cudaGetLastError());
    k_smc_post_init<<<1, 32, 0, ctx->stream>>>(s->B, s->P, s->X, 0);
    SMC_LAUNCHED(s, 1);
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

// the table a sweep reads + one MCMC sweep (+ bookkeeping and the next quantile's first histogram in its last block)
static int smc_enqueue_sweep(kabc_smc *s, bool close_iter) {
    NvtxRange nv("kabc:smc:table+sweep");
    kabc_ctx *ctx = s->ctx;
    const int ci = close_iter ? 1 : 0;
    k_compact<<<s->nblocks_scan, CUT_THREADS, 0, ctx->stream>>>(s->B, s->P, s->X, 0, 1);
    SMC_LAUNCHED(s, 1);
    s->mark();
    const unsigned pb = (unsigned)((s->P.P + 255) / 256);
    switch (s->model.kind) {
    case KABC_MODEL_LV_SSA: {
        k_smc_propose<3><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk);
        s->mark();
        long long cap = (long long)ctx->sm_count * 8;
        const unsigned blocks = (unsigned)((long long)pb < cap ? (long long)pb : cap);
        if (s->model.precision == KABC_F64)
            k_smc_simulate_lv<KABC_F64><<<blocks, 256, 0, ctx->stream>>>(s->B, s->P, s->X, s->model, ctx->rk, ci);
        else
            k_smc_simulate_lv<KABC_F32_ACC64><<<blocks, 256, 0, ctx->stream>>>(s->B, s->P, s->X, s->model, ctx->rk, ci);
        break;
    }
    case KABC_MODEL_GK_OCTILE: {
        k_smc_propose<4><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk);
        s->mark();
        size_t smem;
        int grid = smc_gk_grid(s, smem);
        if (s->model.precision == KABC_F64)
            k_smc_simulate_gk<KABC_F64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->X, s->model, ctx->rk, ci);
        else
            k_smc_simulate_gk<KABC_F32_ACC64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->X, s->model, ctx->rk, ci);
        break;
    }
    default: {
        static const bool split = [] { const char *e = getenv("KABC_SWEEP"); return e && e[0] == 's'; }(); // "split": two kernels
        if (!split) {
            switch (s->model.kind) {
            case KABC_MODEL_NORMAL_MEANSTD: smc_launch_queued_t<KABC_MODEL_NORMAL_MEANSTD>(s, ci); break;
            case KABC_MODEL_MA2_AUTOCOV: smc_launch_queued_t<KABC_MODEL_MA2_AUTOCOV>(s, ci); break;
            case KABC_MODEL_DETERMINISTIC: smc_launch_queued_t<KABC_MODEL_DETERMINISTIC>(s, ci); break;
            default: smc_launch_queued_t<KABC_MODEL_SOCKS>(s, ci); break;
            }
            SMC_LAUNCHED(s, 1);
            s->mark();
            KABC_CUDA_TRY(cudaGetLastError());
            return KABC_OK;
        }
        if (s->P.d <= 2) k_smc_propose<2><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk);
        else if (s->P.d <= 4) k_smc_propose<4><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk);
        else k_smc_propose<KABC_MAX_DIM><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk);
        s->mark();
        switch (s->model.kind) {
        case KABC_MODEL_NORMAL_MEANSTD: smc_launch_list_t<KABC_MODEL_NORMAL_MEANSTD>(s, ci); break;
        case KABC_MODEL_MA2_AUTOCOV: smc_launch_list_t<KABC_MODEL_MA2_AUTOCOV>(s, ci); break;
        case KABC_MODEL_DETERMINISTIC: smc_launch_list_t<KABC_MODEL_DETERMINISTIC>(s, ci); break;
        default: smc_launch_list_t<KABC_MODEL_SOCKS>(s, ci); break;
        }
    }
    }
    SMC_LAUNCHED(s, 2);
    s->mark();
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

static int smc_enqueue_cut(kabc_smc *s) {
    NvtxRange nv("kabc:smc:quantile+cut");
    kabc_ctx *ctx = s->ctx;
    int sel_blocks = (int)((s->P.P + SEL_THREADS * 16 - 1) / (SEL_THREADS * 16));
    if (sel_blocks > ctx->sm_count * 2) sel_blocks = ctx->sm_count * 2;
    if (sel_blocks < 1) sel_blocks = 1;
    s->mark();
    for (int q = 0; q < SEL_INSTANCES; ++q) {
        k_sel<<<sel_blocks, SEL_THREADS, 0, ctx->stream>>>(s->B, s->P, s->X, q == SEL_INSTANCES - 1);
        s->mark();
    }
    k_cut<<<s->nblocks_scan, CUT_THREADS, 0, ctx->stream>>>(s->B, s->P, s->X, s->nblocks_scan);
    s->mark();
    SMC_LAUNCHED(s, SEL_INSTANCES + 1);
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

static int smc_read_ctrl(kabc_smc *s) {
    KABC_CUDA_TRY(cudaMemcpyAsync(s->h_ctrl, s->B.ctrl, sizeof(SmcCtrl), cudaMemcpyDeviceToHost, s->ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    return KABC_OK;
}

static int smc_ctrl_error(kabc_smc *s) {
    switch (s->h_ctrl->err) {
    case 0: return KABC_OK;
    case KABC_ERR_DEGENERATE: return set_error(KABC_ERR_DEGENERATE, "no alive particles left (ESS = 0): cannot resample");
    case KABC_ERR_PEER: return set_error(KABC_ERR_PEER, "a rank of the job did not reach a cross-rank barrier in time");
    case KABC_ERR_STATE: return set_error(KABC_ERR_STATE, "internal: the quantile selection lost track of its candidates");
    default: return set_error(s->h_ctrl->err, "prior sampling failed (truncation too extreme)");
    }
}

// one body of the reference's `while true` loop
static int smc_enqueue_iteration(kabc_smc *s) {
    if (int rc = smc_enqueue_cut(s)) return rc;
    const long long retry_n = 1 + s->P.mcmc_retrys;
    if (retry_n == 1) return smc_enqueue_sweep(s, true);
    for (long long r = 0; r < retry_n; ++r) {
        if (int rc = smc_enqueue_sweep(s, false)) return rc;
        if (r + 1 < retry_n) {
            // ref :192 -- leave the retry loop as soon as enough moves were accepted (the counters are replicated, so
            // every rank takes the same decision)
            if (int rc = smc_read_ctrl(s)) return rc;
            if (s->h_ctrl->err || s->h_ctrl->retry_done) break;
        }
    }
    k_post_iter<<<1, 1, 0, s->ctx->stream>>>(s->B, s->P);
    SMC_LAUNCHED(s, 1);
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

static void smc_drop_graph(kabc_smc *s) {
    if (s->iter_graph) cudaGraphExecDestroy(s->iter_graph);
    s->iter_graph = nullptr;
}

// enqueue one iteration: through the captured graph when the launch sequence is fixed (no retry sweeps, which need the
// host between sweeps), else kernel by kernel
static int smc_launch_iteration(kabc_smc *s) {
    NvtxRange nv("kabc:smc:iteration");
    kabc_ctx *ctx = s->ctx;
    static const bool env_off = [] { const char *e = getenv("KABC_NO_GRAPH"); return e && e[0] == '1'; }();
    const bool eligible = s->graph_ok && !env_off && !s->prof && s->P.mcmc_retrys == 0;
    if (!eligible) return smc_enqueue_iteration(s);
    if (!s->iter_graph) {
        const long long l0 = s->launches, c0 = ctx->launches;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        int rc = KABC_OK;
        if (e == cudaSuccess) {
            rc = smc_enqueue_iteration(s);
            e = cudaStreamEndCapture(ctx->stream, &g);
        }
        s->graph_kernels = (int)(s->launches - l0);
        s->launches = l0; ctx->launches = c0; // nothing ran yet
        if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&s->iter_graph, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess || rc) { // capture not possible here: fall back for good
            cudaGetLastError();
            s->iter_graph = nullptr;
            s->graph_ok = false;
            return smc_enqueue_iteration(s);
        }
    }
    KABC_CUDA_TRY(cudaGraphLaunch(s->iter_graph, ctx->stream));
    SMC_LAUNCHED(s, s->graph_kernels);
    return KABC_OK;
}

extern "C" {

uint64_t kabc_smc_arena_bytes(int64_t nparticles, int d, int world) {
    if (world < 1 || nparticles < 1 || d < 1) return 0;
    return (uint64_t)smc_xlayout((nparticles + world - 1) / world, d, world).bytes + 4096;
}

int kabc_smc_create(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                    const kabc_smc_config_t *cfg, kabc_smc_t **out) {
    if (!ctx || !cfg || !out) return set_error(KABC_ERR_INVALID_ARG, "NULL argument");
    DPriors pri;
    DModel m;
    if (int rc = ingest_priors(prior, d, pri)) return rc;
    if (int rc = smc_check_cfg(cfg, d)) return rc;
    if (int rc = ingest_model(model, d, m)) return rc;
    m.push_mask = push_mask_of(pri);
    const long long N = cfg->nparticles;
    if (ctx->world > 1 && N % ctx->world) return set_error(KABC_ERR_INVALID_ARG, "nparticles must be a multiple of the number of ranks");
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    kabc_smc *s = new kabc_smc();
    const long long Pn = N / ctx->world;
    s->ctx = ctx; s->pri = pri; s->model = m; s->cfg = *cfg;
    s->P.N = N; s->P.P = Pn; s->P.lo = Pn * ctx->rank; s->P.d = d; s->P.TS = table_stride(d);
    s->P.alpha = cfg->alpha; s->P.mcmc_tol = cfg->mcmc_tol; s->P.epstol = cfg->epstol;
    s->P.r_epstol = cfg->r_epstol; s->P.min_r_ess = cfg->min_r_ess; s->P.max_stretch = cfg->max_stretch;
    s->P.sqrt_np = sqrt((double)d);
    s->P.mcmc_retrys = cfg->mcmc_retrys; s->P.max_iterations = cfg->max_iterations;
    s->P.rank = ctx->rank; s->P.world = ctx->world;
    {   // tiles in flight in the queued sweep (k_smc_sweep_q): KABC_PROP_CAP bounds them, 0 = unbounded (the default)
        const char *e = getenv("KABC_PROP_CAP");
        s->P.prop_cap = e ? (unsigned int)strtoul(e, nullptr, 10) : 0u;
    }
    s->X = make_xpeer(ctx);
    s->nblocks_scan = (int)((Pn + SCAN_THREADS - 1) / SCAN_THREADS);
    const size_t nd = (size_t)Pn * d;
    const size_t nd_pad = (nd + 1) & ~(size_t)1, pn_pad = ((size_t)Pn + 1) & ~(size_t)1; // X, lpi 16-byte aligned (double2 loads)
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    // peer-visible block: out of the context's arena (multi rank), else a plain cached buffer
    const XLayout L = smc_xlayout(Pn, d, ctx->world);
    memset(s->B.xb, 0, sizeof s->B.xb);
    if (ctx->world > 1) {
        size_t off = 0;
        if (int rc = arena_alloc(ctx, L.bytes, &off)) { delete s; return rc; }
        s->in_arena = true;
        s->X = make_xpeer(ctx); // the arena may just have been (re)mapped
        for (int r = 0; r < ctx->world; ++r) s->B.xb[r] = (unsigned char *)ctx->arena_map[r] + off;
    } else {
        A(s->xlocal.alloc(ctx, L.bytes));
        s->B.xb[0] = s->xlocal.p;
    }
    s->B.o_th = (long long)L.o_th; s->B.o_alive = (long long)L.o_alive;
    A(s->state.alloc(ctx, nd_pad + 2 * pn_pad));
    A(s->thp.alloc(ctx, nd)); A(s->lpip.alloc(ctx, Pn)); A(s->work.alloc(ctx, Pn)); A(s->fill.alloc(ctx, (size_t)(Pn + 255) / 256 + 2));
    A(s->alive.alloc(ctx, Pn)); A(s->blockcnt.alloc(ctx, s->nblocks_scan));
    A(s->hist.alloc(ctx, SEL_BINS)); A(s->cand.alloc(ctx, SEL_CAP)); A(s->ctrl.alloc(ctx, 1));
    const long long log_cap = 1 << 14;
    A(s->log.alloc(ctx, log_cap));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&s->h_ctrl, sizeof(SmcCtrl));
    if (e != cudaSuccess) {
        if (s->in_arena) arena_release(ctx);
        delete s;
        return set_error(KABC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    memset(s->h_ctrl, 0, sizeof(SmcCtrl));
    s->B.th = s->state.p; s->B.X = s->state.p + nd_pad; s->B.lpi = s->B.X + pn_pad;
    s->B.alive = s->alive.p; s->B.thp = s->thp.p; s->B.lpip = s->lpip.p; s->B.work = s->work.p;
    s->B.fill = s->fill.p;
    s->B.blockcnt = s->blockcnt.p; s->B.hist = s->hist.p; s->B.cand = s->cand.p; s->B.ctrl = s->ctrl.p;
    s->B.log = s->log.p; s->B.log_cap = log_cap;
    memset(&s->B.tr, 0, sizeof s->B.tr);
    s->B.trace_on = 0;
    cudaError_t e2 = cudaMemsetAsync(s->ctrl.p, 0, sizeof(SmcCtrl), ctx->stream);
    if (e2 == cudaSuccess) e2 = cudaMemsetAsync(s->hist.p, 0, sizeof(unsigned int) * SEL_BINS, ctx->stream);
    // launch geometry of the sweep
    if (e2 == cudaSuccess && m.kind == KABC_MODEL_GK_OCTILE) {
        const int smem = (int)gk_smem_bytes(m.n_draws, m.precision);
        if (m.precision == KABC_F64) {
            e2 = cudaFuncSetAttribute(k_smc_init_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_smc_simulate_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        } else {
            e2 = cudaFuncSetAttribute(k_smc_init_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(k_smc_simulate_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        }
    }
    if (e2 != cudaSuccess) {
        kabc_smc_destroy(s);
        return set_error(KABC_ERR_CUDA, "smc handle setup failed: %s", cudaGetErrorString(e2));
    }
    *out = s;
    return KABC_OK;
}

int kabc_smc_destroy(kabc_smc_t *s) {
    if (!s) return KABC_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    smc_drop_graph(s);
    if (s->in_arena) arena_release(s->ctx);
    if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
    delete s;
    return KABC_OK;
}

int kabc_smc_init(kabc_smc_t *s) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (int rc = smc_enqueue_init(s)) return rc;
    if (int rc = smc_read_ctrl(s)) return rc;
    if (int rc = smc_ctrl_error(s)) return rc;
    s->inited = true;
    return KABC_OK;
}

int kabc_smc_iterate(kabc_smc_t *s, int *stop) {
    if (!s || !stop) return set_error(KABC_ERR_INVALID_ARG, "NULL argument");
    if (!s->inited) return set_error(KABC_ERR_STATE, "kabc_smc_init must be called first");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (int rc = smc_launch_iteration(s)) return rc;
    if (int rc = smc_read_ctrl(s)) return rc;
    if (int rc = smc_ctrl_error(s)) return rc;
    *stop = s->h_ctrl->stop;
    return KABC_OK;
}

int kabc_smc_iterate_n(kabc_smc_t *s, int n, int ignore_stop, int *done, float *out_ms) {
    if (!s || n < 0) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    if (!s->inited) return set_error(KABC_ERR_STATE, "kabc_smc_init must be called first");
    kabc_ctx *ctx = s->ctx;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    KABC_CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    int it = 0;
    for (; it < n; ++it) {
        if (int rc = smc_launch_iteration(s)) return rc;
        if (!ignore_stop) {
            if (int rc = smc_read_ctrl(s)) return rc;
            if (int rc = smc_ctrl_error(s)) return rc;
            if (s->h_ctrl->stop) { ++it; break; }
        }
    }
    KABC_CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    if (int rc = smc_read_ctrl(s)) return rc;
    if (int rc = smc_ctrl_error(s)) return rc;
    if (done) *done = it;
    if (out_ms) KABC_CUDA_TRY(cudaEventElapsedTime(out_ms, ctx->ev0, ctx->ev1));
    return KABC_OK;
}

// n iterations, each preceded by an L2 flush (a memset of flush_bytes, outside the timed region), all enqueued without a
// host round trip; out_ms[i] = device time of iteration i between two events on the launching stream
int kabc_smc_bench_steps(kabc_smc_t *s, int n, uint64_t flush_bytes, float *out_ms) {
    if (!s || n < 0 || !out_ms) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    if (!s->inited) return set_error(KABC_ERR_STATE, "kabc_smc_init must be called first");
    kabc_ctx *ctx = s->ctx;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf<unsigned char> flush;
    if (flush_bytes) KABC_CUDA_TRY(flush.alloc(ctx, (size_t)flush_bytes));
    std::vector<cudaEvent_t> ev(2 * (size_t)n);
    for (auto &e : ev) KABC_CUDA_TRY(cudaEventCreate(&e));
    int rc = KABC_OK;
    for (int it = 0; it < n && !rc; ++it) {
        if (flush_bytes) KABC_CUDA_TRY(cudaMemsetAsync(flush.p, it & 0xff, (size_t)flush_bytes, ctx->stream));
        KABC_CUDA_TRY(cudaEventRecord(ev[2 * it], ctx->stream));
        rc = smc_launch_iteration(s);
        KABC_CUDA_TRY(cudaEventRecord(ev[2 * it + 1], ctx->stream));
    }
    if (!rc) rc = smc_read_ctrl(s);
    if (!rc) rc = smc_ctrl_error(s);
    for (int it = 0; it < n; ++it) {
        out_ms[it] = 0.f;
        if (!rc) cudaEventElapsedTime(&out_ms[it], ev[2 * it], ev[2 * it + 1]);
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

int kabc_smc_profile_iteration(kabc_smc_t *s, float *out_us, int cap, int *out_n) {
    if (!s || !out_us || !out_n) return set_error(KABC_ERR_INVALID_ARG, "NULL argument");
    if (!s->inited) return set_error(KABC_ERR_STATE, "kabc_smc_init must be called first");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    std::vector<cudaEvent_t> evs;
    s->prof = &evs;
    int rc = smc_enqueue_iteration(s);
    s->prof = nullptr;
    if (!rc) rc = smc_read_ctrl(s);
    int n = 0;
    for (size_t q = 1; q < evs.size(); ++q) {
        float ms = 0;
        cudaEventElapsedTime(&ms, evs[q - 1], evs[q]);
        if (n < cap) out_us[n++] = ms * 1e3f;
    }
    for (cudaEvent_t e : evs) cudaEventDestroy(e);
    *out_n = n;
    if (rc) return rc;
    return smc_ctrl_error(s);
}

int kabc_smc_get_state(kabc_smc_t *s, double *theta, double *X, double *lpi, uint8_t *alive) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    kabc_ctx *ctx = s->ctx;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t N = (size_t)s->P.N;
    const int d = s->P.d;
    cudaStream_t st = ctx->stream;
    if (ctx->world == 1) {
        if (theta) KABC_CUDA_TRY(cudaMemcpyAsync(theta, s->B.th, 8 * N * d, cudaMemcpyDeviceToHost, st));
        if (X) KABC_CUDA_TRY(cudaMemcpyAsync(X, s->B.X, 8 * N, cudaMemcpyDeviceToHost, st));
        if (lpi) KABC_CUDA_TRY(cudaMemcpyAsync(lpi, s->B.lpi, 8 * N, cudaMemcpyDeviceToHost, st));
        if (alive) KABC_CUDA_TRY(cudaMemcpyAsync(alive, s->B.alive, N, cudaMemcpyDeviceToHost, st));
        KABC_CUDA_TRY(cudaStreamSynchronize(st));
        return KABC_OK;
    }
    // multi rank (collective: every rank calls this): every rank publishes its rows in its table, then reads all tables
    DevBuf<double> g_th, g_X, g_lpi;
    DevBuf<unsigned char> g_alive;
    if (theta) KABC_CUDA_TRY(g_th.alloc(ctx, N * d));
    if (X) KABC_CUDA_TRY(g_X.alloc(ctx, N));
    if (lpi) KABC_CUDA_TRY(g_lpi.alloc(ctx, N));
    if (alive) KABC_CUDA_TRY(g_alive.alloc(ctx, N));
    k_compact<<<s->nblocks_scan, CUT_THREADS, 0, st>>>(s->B, s->P, s->X, 1, 1);
    k_gather_full<<<ctx->sm_count * 4, 256, 0, st>>>(s->B, s->P, s->X, g_th.p, g_X.p, g_lpi.p, g_alive.p);
    SMC_LAUNCHED(s, 2);
    KABC_CUDA_TRY(cudaGetLastError());
    if (theta) KABC_CUDA_TRY(cudaMemcpyAsync(theta, g_th.p, 8 * N * d, cudaMemcpyDeviceToHost, st));
    if (X) KABC_CUDA_TRY(cudaMemcpyAsync(X, g_X.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (lpi) KABC_CUDA_TRY(cudaMemcpyAsync(lpi, g_lpi.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (alive) KABC_CUDA_TRY(cudaMemcpyAsync(alive, g_alive.p, N, cudaMemcpyDeviceToHost, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    if (int rc = smc_read_ctrl(s)) return rc;
    return smc_ctrl_error(s);
}

int kabc_smc_set_state(kabc_smc_t *s, const double *theta, const double *X, const double *lpi, const uint8_t *alive) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const size_t N = (size_t)s->P.N, Pn = (size_t)s->P.P, lo = (size_t)s->P.lo;
    cudaStream_t st = s->ctx->stream;
    if (theta)
        for (int k = 0; k < s->P.d; ++k)
            KABC_CUDA_TRY(cudaMemcpyAsync(s->B.th + (size_t)k * Pn, theta + (size_t)k * N + lo, 8 * Pn, cudaMemcpyHostToDevice, st));
    if (X) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.X, X + lo, 8 * Pn, cudaMemcpyHostToDevice, st));
    if (lpi) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.lpi, lpi + lo, 8 * Pn, cudaMemcpyHostToDevice, st));
    if (alive) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.alive, alive + lo, Pn, cudaMemcpyHostToDevice, st));
    k_recount<<<1, 1024, 0, st>>>(s->B, s->P);
    k_smc_post_init<<<1, 32, 0, st>>>(s->B, s->P, s->X, 1);
    SMC_LAUNCHED(s, 2);
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    return KABC_OK;
}

int kabc_smc_get_scalars(kabc_smc_t *s, double *eps, int32_t *flag, int64_t *iteration, int64_t *n_alive,
                         int64_t *accepted, int64_t *cost_evals, int64_t *next_epoch, int64_t *events) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    if (int rc = smc_read_ctrl(s)) return rc;
    const SmcCtrl *c = s->h_ctrl;
    if (eps) *eps = c->eps;
    if (flag) *flag = c->flag;
    if (iteration) *iteration = c->iteration;
    if (n_alive) *n_alive = c->ess; /* ESS after the last cut, what the reference prints (src/smc.jl:142-143) */
    if (accepted) *accepted = (int64_t)c->accepted;
    if (cost_evals) *cost_evals = (int64_t)c->cost_evals;
    if (next_epoch) *next_epoch = c->epoch;
    if (events) *events = (int64_t)c->events;
    return KABC_OK;
}

int64_t kabc_smc_get_log(kabc_smc_t *s, kabc_smc_log_t *log, int64_t cap) {
    if (!s) return -1;
    cudaSetDevice(s->ctx->device);
    if (smc_read_ctrl(s)) return -1;
    int64_t n = s->h_ctrl->iteration;
    if (n > s->B.log_cap) n = s->B.log_cap;
    int64_t m = n < cap ? n : cap;
    if (log && m > 0) {
        if (cudaMemcpyAsync(log, s->B.log, sizeof(kabc_smc_log_t) * (size_t)m, cudaMemcpyDeviceToHost, s->ctx->stream) != cudaSuccess) return -1;
        cudaStreamSynchronize(s->ctx->stream);
    }
    return n;
}

int64_t kabc_smc_kernel_launches(kabc_smc_t *s) { return s ? s->launches : -1; }

int kabc_smc_trace_enable(kabc_smc_t *s, int on) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    if (on && s->ctx->world > 1) return set_error(KABC_ERR_STATE, "the replay trace is a single-rank facility");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    KABC_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    if (on && !s->ta.p) {
        const size_t N = (size_t)s->P.N;
        KABC_CUDA_TRY(s->ta.alloc(N)); KABC_CUDA_TRY(s->tb.alloc(N)); KABC_CUDA_TRY(s->tz.alloc(N));
        KABC_CUDA_TRY(s->tlprob.alloc(N)); KABC_CUDA_TRY(s->tlpip.alloc(N)); KABC_CUDA_TRY(s->txp.alloc(N));
        KABC_CUDA_TRY(s->tdec.alloc(N));
        if (!s->thp.p) { KABC_CUDA_TRY(s->thp.alloc(s->ctx, N * s->P.d)); s->B.thp = s->thp.p; }
        s->B.tr.a = s->ta.p; s->B.tr.b = s->tb.p; s->B.tr.z = s->tz.p; s->B.tr.lprob = s->tlprob.p;
        s->B.tr.lpip = s->tlpip.p; s->B.tr.xp = s->txp.p; s->B.tr.dec = s->tdec.p;
    }
    s->B.trace_on = on ? 1 : 0;
    smc_drop_graph(s); // the captured launches hold SmcBufs by value
    return KABC_OK;
}

int kabc_smc_get_trace(kabc_smc_t *s, int64_t *a, int64_t *b, double *z, double *lprob, double *lpi_p, double *xp,
                       uint8_t *decision, double *theta_p) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    if (!s->ta.p) return set_error(KABC_ERR_STATE, "trace was never enabled");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const size_t N = (size_t)s->P.N;
    cudaStream_t st = s->ctx->stream;
    if (a) KABC_CUDA_TRY(cudaMemcpyAsync(a, s->ta.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (b) KABC_CUDA_TRY(cudaMemcpyAsync(b, s->tb.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (z) KABC_CUDA_TRY(cudaMemcpyAsync(z, s->tz.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (lprob) KABC_CUDA_TRY(cudaMemcpyAsync(lprob, s->tlprob.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (lpi_p) KABC_CUDA_TRY(cudaMemcpyAsync(lpi_p, s->tlpip.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (xp) KABC_CUDA_TRY(cudaMemcpyAsync(xp, s->txp.p, 8 * N, cudaMemcpyDeviceToHost, st));
    if (decision) KABC_CUDA_TRY(cudaMemcpyAsync(decision, s->tdec.p, N, cudaMemcpyDeviceToHost, st));
    if (theta_p) KABC_CUDA_TRY(cudaMemcpyAsync(theta_p, s->thp.p, 8 * N * s->P.d, cudaMemcpyDeviceToHost, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    return KABC_OK;
}

int kabc_smc_run(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const kabc_model_t *model,
                 const kabc_smc_config_t *cfg, double *out_theta, uint8_t *out_alive, double *out_cost, double *out_eps,
                 int64_t *out_iterations, int64_t *out_cost_evals, kabc_smc_log_t *log, int64_t log_cap) {
    kabc_smc *s = nullptr;
    const bool dbg = getenv("KABC_DEBUG_TIMING") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    const double t0 = now();
    if (int rc = kabc_smc_create(ctx, prior, d, model, cfg, &s)) return rc;
    const double t1 = now();
    int rc = kabc_smc_init(s);
    const double t2 = now();
    if (!rc && s->P.mcmc_retrys > 0) {
        // the retry loop needs the host between sweeps anyway (ref :192): plain stepping
        int stop = 0;
        while (!rc && !stop) rc = kabc_smc_iterate(s, &stop);
    } else if (!rc) {
        // One iteration is always queued AHEAD of the one whose `stop` flag the host is waiting for, so kernel
        // launches and the flag read-back overlap with device work.  The look-ahead iteration does nothing on the
        // device when `stop` was set (smc_skip).
        cudaStream_t st = ctx->stream;
        SmcCtrl *slots = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaError_t e = cudaMallocHost((void **)&slots, 2 * sizeof(SmcCtrl));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        if (e != cudaSuccess) rc = set_error(KABC_ERR_CUDA, "pipeline setup failed: %s", cudaGetErrorString(e));
        if (!rc) {
            k_set_honor_stop<<<1, 1, 0, st>>>(s->B, 1);
            SMC_LAUNCHED(s, 1);
        }
        auto enqueue = [&](int slot) -> int {
            if (int r2 = smc_launch_iteration(s)) return r2;
            if (cudaMemcpyAsync(&slots[slot], s->B.ctrl, sizeof(SmcCtrl), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaEventRecord(ev[slot], st) != cudaSuccess)
                return set_error(KABC_ERR_CUDA, "pipeline enqueue failed");
            return KABC_OK;
        };
        if (!rc) rc = enqueue(0);
        for (int k = 0; !rc; ++k) {
            rc = enqueue((k + 1) & 1);
            if (rc) break;
            if (cudaEventSynchronize(ev[k & 1]) != cudaSuccess) { rc = set_error(KABC_ERR_CUDA, "event wait failed"); break; }
            const SmcCtrl &c = slots[k & 1];
            if (c.err || c.stop) {
                cudaEventSynchronize(ev[(k + 1) & 1]); // the look-ahead iteration was skipped on the device
                break;
            }
        }
        if (!rc) rc = smc_read_ctrl(s);
        if (!rc) rc = smc_ctrl_error(s);
        if (ev[0]) cudaEventDestroy(ev[0]);
        if (ev[1]) cudaEventDestroy(ev[1]);
        if (slots) cudaFreeHost(slots);
    }
    const double t3 = now();
    double eps_out = 0;
    long long it_out = 0, evals_out = 0;
    if (!rc) { eps_out = s->h_ctrl->eps; it_out = s->h_ctrl->iteration; evals_out = (long long)s->h_ctrl->cost_evals; }
    if (!rc) rc = kabc_smc_get_state(s, out_theta, out_cost, nullptr, out_alive);
    if (!rc && out_theta) { // ref src/smc.jl:200: the returned particles are push_p(prior, .)
        const long long N = s->P.N;
        for (int k = 0; k < d; ++k)
            if (prior_is_discrete(s->pri.p[k]))
                for (long long i = 0; i < N; ++i) out_theta[(long long)k * N + i] = nearbyint(out_theta[(long long)k * N + i]);
    }
    if (!rc) {
        if (out_eps) *out_eps = eps_out;
        if (out_iterations) *out_iterations = it_out;
        if (out_cost_evals) *out_cost_evals = evals_out;
        if (log && log_cap > 0 && kabc_smc_get_log(s, log, log_cap) < 0) rc = set_error(KABC_ERR_CUDA, "log copy failed");
    }
    const double t4 = now();
    std::string keep = g_last_error;
    kabc_smc_destroy(s);
    g_last_error = keep;
    if (dbg)
        fprintf(stderr, "[kabc_smc_run] create %.1f ms, init %.1f ms, iterate %.1f ms, copy-out %.1f ms, destroy %.1f ms\n",
                1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (now() - t4));
    return rc;
}

} // extern "C"
