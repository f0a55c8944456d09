// kabc_smc.cu -- smc(prior, cost; ...) on device.  Restates src/smc.jl:92-206 of KissABC.jl 3.0.1:
//   init                      :119-129   k_smc_init / k_smc_init_prior + k_smc_init_gk
//   eps = quantile(Xs[alive]) :134       k_sel_hist<0>, k_sel_hist<1>, k_sel_final  (exact two-rank bucket select on
//                                        FP64 keys + Statistics.jl type-7 interpolation)
//   alive cut, flag, ESS      :135-142   k_alive_cut
//   cyclic-tiling resample    :145-153   k_alive_cut (decision + scan), k_resample_scatter, k_resample_gather
//   propose                   :160-167   k_smc_propose  (+ prior-MH pre-test :172-175, builds the work list)
//   simulate + accept         :176-189   k_smc_simulate / k_smc_simulate_gk
//   retry / stop rules        :156-159,192-198   post_sweep()/post_iter() run by the LAST block of the sweep kernel
// Every scalar that steers control flow lives in SmcCtrl in device memory; the host only reads `stop`.
// State is SoA FP64: th[k*N+i], X[i], lpi[i], alive[i]; two copies (ping-pong): every iteration gathers
// (resample) or copies (no resample) into the other copy, so the host always knows which copy is current.
//
// Multi-GPU (one process per GPU): the state is replicated, the sweep is sharded.  Rank r proposes/simulates/
// accepts particles [r N/G, (r+1) N/G); the updated shard rows and four counters are all-gathered (NCCL over
// NVLink) after each sweep; quantile, cut and resample are computed redundantly from identical replicas, so no
// scalar ever needs a broadcast.  Philox counters are keyed by the GLOBAL particle id: results are bit-identical
// for any G.
#include <ctime>
#include "kabc_host.hpp"
#include "kabc_gk.cuh"
#include "kabc_nccl.hpp"

namespace kabc {

constexpr int SEL_LOG2_BINS = 12;
constexpr int SEL_BINS = 1 << SEL_LOG2_BINS;
constexpr int SEL_CAP = 4096; // candidates sorted exactly by one block
constexpr int SEL_THREADS = 512;
constexpr int SCAN_THREADS = 1024; // particles per block of the cut / scatter kernels

struct RankPartial { // what a rank contributes to the sweep bookkeeping
    unsigned long long accepted, work, events, minkey;
    unsigned long long pushed; // records this rank left in every peer's inbox during the sweep (packed pushes)
};

struct SmcCtrl {
    double eps, eps_prev, xmin, gamma;
    unsigned long long xmin_key; // running minimum over the alive costs (as an ordered key)
    unsigned long long klo, khi, v0key, v1key;
    long long below, cnt, r0, r1; // bucket-select bookkeeping
    unsigned long long sel_below;  // alive keys under the guessed lower bound of pass 0
    long long n_alive;            // number of alive particles (input of the next quantile)
    long long ess;                // ESS = sum(alive) right after the cut (what the reference prints)
    unsigned long long accepted, cost_evals, events;
    unsigned long long sw_accepted, sw_events, sw_minkey; // per-sweep partials of this rank
    long long iteration;
    unsigned int work_count, cand_count, epoch, lv_head;
    unsigned int push_count; // rows k_smc_propose finalised and packed into the peers' inboxes this sweep (multi GPU)
    unsigned int tk_hist, tk_final, tk_cut, tk_gather, tk_sim;
    int flag, resample, stop, cur, err, sweeps, retry_done, resampled_log, sel_done, bounds_known;
    int honor_stop; // kabc_smc_run enqueues one iteration ahead: once `stop` is set the queued kernels do nothing
};

struct SmcParams { // launch constants
    long long N;
    int d;
    double alpha, mcmc_tol, epstol, r_epstol, min_r_ess, max_stretch;
    long long mcmc_retrys;
    int max_iterations;
    int rank, world;
};

struct SmcTrace {
    long long *a, *b;
    double *z, *lprob, *lpip, *xp, *thp;
    unsigned char *dec;
};

constexpr int KABC_MAX_PEERS = 16;

struct SmcBufs {
    // the six state arrays live in ONE slab per rank: copy c at slab + c*(d+2)*N = [th (d*N) | X (N) | lpi (N)]
    double *th[2], *X[2], *lpi[2];
    // peer-memory replicas (multi GPU): peer[r] is rank r's slab mapped into this process (cudaIpc over NVLink);
    // n_peers == 0 -> no direct pushes (single GPU, or the NCCL all-gather fallback)
    double *peer[KABC_MAX_PEERS];
    int n_peers;
    int shard_rows; // 1: only X is replicated; theta/lpi rows are read from their owner's slab in k_smc_propose
    // packed pushes: rows that are final after k_smc_propose are few and scattered, and as single 8-byte NVLink stores
    // they cost one packet each.  They are written instead as dense records into an inbox inside every peer's slab
    // (planes [index | th_0..th_{d-1} | X | lpi], one region per (sweep parity, source rank)), which k_apply_inbox
    // scatters locally after the sweep barrier.
    int packed;
    long long inbox_off; // offset (doubles) of the inbox region inside a slab
    unsigned char *alive;
    double *thp, *lpip;
    unsigned int *work, *idxalive, *blockcnt, *hist;
    unsigned long long *cand;
    SmcCtrl *ctrl;
    RankPartial *partial; // [world]
    kabc_smc_log_t *log;
    long long log_cap;
    SmcTrace tr;
    int trace_on;
};

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
// true for exactly one block: the last one to arrive (its reads see every other block's writes)
__device__ __forceinline__ bool last_block(unsigned int *ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// ------------------------------------------------------------------ init, ref src/smc.jl:119-129
template <int KIND, int PREC>
__global__ void __launch_bounds__(256)
k_smc_init(SmcBufs B, SmcParams P, DPriors pri, DModel m, RoundKeys rk, long long lo, long long hi) {
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = P.N;
    unsigned long long key = ~0ull, ev_u = 0;
    if (i < hi) {
        Stream st(rk, ST_PRIOR, (uint32_t)i, 0u);
        bool ok = true;
#pragma unroll 1
        for (int k = 0; k < P.d; ++k) {
            double x;
            ok &= prior1_sample(pri.p[k], st, x);
            B.th[0][(long long)k * N + i] = x;
        }
        if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
        long long ev;
        double *thv = B.th[0];
        double X = cost_thread<KIND, PREC>(m, rk, ST_COST_INIT, (uint32_t)i, 0u, [&](int k) { return thv[(long long)k * N + i]; }, ev);
        B.X[0][i] = X;
        B.lpi[0][i] = prior_logpdf_pushed(pri, [&](int k) { return thv[(long long)k * N + i]; });
        B.alive[i] = 1;
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

__global__ void k_smc_init_prior(SmcBufs B, SmcParams P, DPriors pri, RoundKeys rk, long long lo, long long hi) {
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const long long N = P.N;
    Stream st(rk, ST_PRIOR, (uint32_t)i, 0u);
    bool ok = true;
    for (int k = 0; k < P.d; ++k) {
        double x;
        ok &= prior1_sample(pri.p[k], st, x);
        B.th[0][(long long)k * N + i] = x;
    }
    if (!ok) B.ctrl->err = KABC_ERR_INVALID_ARG;
    double *thv = B.th[0];
    B.lpi[0][i] = prior_logpdf_pushed(pri, [&](int k) { return thv[(long long)k * N + i]; });
    B.alive[i] = 1;
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS)
k_smc_init_gk(SmcBufs B, SmcParams P, DModel m, RoundKeys rk, long long lo, long long hi) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    const long long N = P.N;
    for (long long i = lo + blockIdx.x; i < hi; i += gridDim.x) {
        const double *th = B.th[0];
        double c = cost_gk_block<PREC>(m, rk, ST_COST_INIT, (uint32_t)i, 0u, pushk(m, 0, th[i]), pushk(m, 1, th[N + i]),
                                    pushk(m, 2, th[2 * N + i]), pushk(m, 3, th[3 * N + i]), gk_smem);
        if (threadIdx.x == 0) {
            B.X[0][i] = c;
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
        c->cost_evals = (unsigned long long)P.N;
    }
}

// after the init kernels: fold this rank's partials (single GPU) or everybody's (after the all-gather)
__global__ void k_smc_post_init(SmcBufs B, SmcParams P, int from_partials) {
    SmcCtrl *c = B.ctrl;
    if (from_partials) {
        unsigned long long mk = ~0ull, ev = 0;
        for (int r = 0; r < P.world; ++r) {
            mk = B.partial[r].minkey < mk ? B.partial[r].minkey : mk;
            ev += B.partial[r].events;
        }
        c->xmin_key = mk; c->events = ev;
    } else {
        c->xmin_key = c->sw_minkey; c->events = c->sw_events;
    }
    c->sw_minkey = ~0ull; c->sw_events = 0; c->sw_accepted = 0;
}
__global__ void k_smc_write_partial(SmcBufs B, SmcParams P) {
    SmcCtrl *c = B.ctrl;
    RankPartial p;
    p.accepted = c->sw_accepted; p.work = c->work_count; p.events = c->sw_events; p.minkey = c->sw_minkey;
    p.pushed = 0;
    B.partial[P.rank] = p;
}

// ------------------------------------------------------------------ quantile, ref src/smc.jl:134 (Statistics type 7)
// Exact selection of the two adjacent order statistics v[j], v[j+1] by range narrowing: histogram the alive keys
// inside [klo,khi] into 4096 equal-width key bins, keep the bins holding the two ranks, repeat; once at most
// SEL_CAP keys remain they are compacted and sorted by one block.  Keys are the order-preserving u64 image of the
// doubles, so bins, ranks and the result are exact (no floating point in the selection).
struct SelRange {
    unsigned long long klo, khi;
    long long below, cnt, r0, r1;
    int shift;
};
__device__ __forceinline__ int sel_shift(unsigned long long klo, unsigned long long khi) {
    const unsigned long long range = khi - klo;
    const int bl = range ? 64 - __clzll((long long)range) : 0;
    return bl > SEL_LOG2_BINS ? bl - SEL_LOG2_BINS : 0;
}
// block-wide (SEL_THREADS threads): scan SEL_BINS counters, narrow the range to the bins holding ranks r0 and r1
__device__ void sel_scan_narrow(const unsigned int *hist, bool hist_is_global, SelRange &R, unsigned int *s_scan,
                                unsigned long long *s_res) {
    constexpr int PER = SEL_BINS / SEL_THREADS;
    unsigned int loc[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int bin = threadIdx.x * PER + q;
        loc[q] = hist_is_global ? __ldcg(&hist[bin]) : hist[bin];
        sum += loc[q];
    }
    s_scan[threadIdx.x] = sum;
    if (threadIdx.x < 4) s_res[threadIdx.x] = 0;
    __syncthreads();
    for (int o = 1; o < SEL_THREADS; o <<= 1) {
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

// 4 consecutive particles per thread: one uchar4 + two double2 loads
__device__ __forceinline__ void load4(const unsigned char *alive, const double *X, long long base, long long N,
                                      unsigned int (&a)[4], double (&x)[4]) {
    if (base + 3 < N && (reinterpret_cast<unsigned long long>(X) & 15ull) == 0) {
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

template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_hist(SmcBufs B, SmcParams P) {
    __shared__ unsigned int sh[SEL_BINS];
    __shared__ unsigned int s_scan[SEL_THREADS];
    __shared__ unsigned long long s_res[4];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    SelRange R;
    double gamma = 0.0;
    unsigned long long klo_true = 0;
    if (PASS == 0) {
        const long long n = c->n_alive;
        if (n <= 0) {
            if (blockIdx.x == 0 && threadIdx.x == 0) c->err = KABC_ERR_DEGENERATE;
            return;
        }
        // aleph = n*p + (1-p); j = clamp(trunc(aleph),1,n-1); gamma = clamp(aleph-j,0,1)
        const double aleph = xadd(xmul((double)n, P.alpha), xsub(1.0, P.alpha));
        long long j = (long long)aleph;
        if (j > n - 1) j = n - 1;
        if (j < 1) j = 1;
        gamma = xsub(aleph, (double)j);
        gamma = gamma < 0.0 ? 0.0 : (gamma > 1.0 ? 1.0 : gamma);
        R.r0 = (n == 1) ? 0 : j - 1; // 0-based rank of v[j]
        R.r1 = (n == 1) ? 0 : j;     // 0-based rank of v[j+1]
        klo_true = c->xmin_key;
        R.klo = klo_true;
        R.khi = ~0ull;
        if (c->bounds_known && dfinite(c->eps)) {
            R.khi = dkey(c->eps);
            // guess: the new quantile sits in the top binade of the alive costs (it does for any alpha that is not
            // tiny); keys under the guess are only counted.  A wrong guess is detected below and costs one more pass.
            const unsigned long long g = dkey(xmul(c->eps, 0.5));
            if (c->eps > 0.0 && g > R.klo && g < R.khi) R.klo = g;
        }
        if (R.khi < R.klo) { R.khi = ~0ull; R.klo = klo_true; }
        R.below = 0; R.cnt = n;
    } else {
        if (c->sel_done || c->cnt <= SEL_CAP) return;
        R.klo = c->klo; R.khi = c->khi; R.below = c->below; R.cnt = c->cnt; R.r0 = c->r0; R.r1 = c->r1;
    }
    R.shift = sel_shift(R.klo, R.khi);
    for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) sh[q] = 0;
    __syncthreads();
    const double *X = B.X[c->cur];
    unsigned long long nbelow = 0;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long base = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < P.N; base += stride) {
        unsigned int a[4];
        double x[4];
        load4(B.alive, X, base, P.N, a, x);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (!a[q]) continue;
            const unsigned long long key = dkey(x[q]);
            if (key >= R.klo && key <= R.khi) atomicAdd(&sh[(key - R.klo) >> R.shift], 1u);
            else if (PASS == 0 && key < R.klo) nbelow += 1;
        }
    }
    if (PASS == 0) {
        nbelow = warp_sum_u64(nbelow);
        if ((threadIdx.x & 31) == 0 && nbelow) atomicAdd(&c->sel_below, nbelow);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) {
        const unsigned int v = sh[q];
        if (v) atomicAdd(&B.hist[q], v);
    }
    if (!last_block(&c->tk_hist)) return;
    bool guess_failed = false;
    if (PASS == 0) {
        const long long below = (long long)c->sel_below;
        if (R.r0 < below) { // a rank lies under the guessed bound: hand the whole range [true min, khi] to the next pass
            guess_failed = true;
            R.klo = klo_true; R.below = 0; R.cnt = c->n_alive; R.shift = 1;
        } else {
            R.below = below;
        }
    }
    if (!guess_failed) sel_scan_narrow(B.hist, true, R, s_scan, s_res);
    for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) B.hist[q] = 0;
    if (threadIdx.x == 0) {
        c->tk_hist = 0;
        c->sel_below = 0;
        c->klo = R.klo; c->khi = R.khi; c->below = R.below; c->cnt = R.cnt; c->r0 = R.r0; c->r1 = R.r1;
        c->sel_done = (R.shift == 0); // bins were single keys: klo/khi ARE v[j], v[j+1]
        c->v0key = R.klo; c->v1key = R.khi;
        if (PASS == 0) { // opens the iteration, ref :132-133
            c->gamma = gamma;
            c->iteration += 1;
            c->eps_prev = c->eps;
            c->sweeps = 0; c->retry_done = 0; c->accepted = 0; c->resampled_log = 0;
        }
    }
}

__global__ void __launch_bounds__(SEL_THREADS) k_sel_final(SmcBufs B, SmcParams P) {
    __shared__ unsigned long long s_keys[SEL_CAP]; // candidates (sort path) or histogram (slow path)
    __shared__ unsigned int s_scan[SEL_THREADS];
    __shared__ unsigned long long s_res[4];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    const double *X = B.X[c->cur];
    const bool compact = !c->sel_done && c->cnt <= SEL_CAP;
    if (compact) {
        const unsigned long long klo = c->klo, khi = c->khi;
        const long long stride = (long long)gridDim.x * blockDim.x * 4;
        const long long nloop = (P.N + stride - 1) / stride;
        for (long long it = 0; it < nloop; ++it) {
            const long long base = it * stride + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
            unsigned int a[4] = {0u, 0u, 0u, 0u};
            double x[4] = {0.0, 0.0, 0.0, 0.0};
            if (base < P.N) load4(B.alive, X, base, P.N, a, x);
            unsigned long long keys[4];
            bool ins[4], any = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                keys[q] = dkey(x[q]);
                ins[q] = a[q] && keys[q] >= klo && keys[q] <= khi;
                any |= ins[q];
            }
            if (__ballot_sync(0xffffffffu, any) == 0) continue; // candidates are rare (<= 4096 of N)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned long long key = keys[q];
                const bool in = ins[q];
                const unsigned int ball = __ballot_sync(0xffffffffu, in);
                if (ball) {
                    const unsigned int lane = threadIdx.x & 31;
                    unsigned int pos = 0;
                    if (lane == 0) pos = atomicAdd(&c->cand_count, (unsigned int)__popc(ball));
                    pos = __shfl_sync(0xffffffffu, pos, 0);
                    if (in) B.cand[pos + __popc(ball & ((1u << lane) - 1u))] = key;
                }
            }
        }
    }
    if (!last_block(&c->tk_final)) return;
    unsigned long long v0, v1;
    if (c->sel_done) {
        v0 = c->v0key; v1 = c->v1key;
    } else if (compact) {
        const int n = (int)c->cnt;
        int npad = 2;
        while (npad < n) npad <<= 1;
        for (int q = threadIdx.x; q < npad; q += blockDim.x) s_keys[q] = q < n ? __ldcg(&B.cand[q]) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= npad; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                    const bool up = (i & k) == 0;
                    const unsigned long long a = s_keys[i], b = s_keys[l];
                    if ((a > b) == up) { s_keys[i] = b; s_keys[l] = a; }
                }
                __syncthreads();
            }
        v0 = s_keys[c->r0 - c->below];
        v1 = s_keys[c->r1 - c->below];
    } else {
        // rare slow path (massive ties / pathological spread): this block alone keeps narrowing over all N
        SelRange R;
        R.klo = c->klo; R.khi = c->khi; R.below = c->below; R.cnt = c->cnt; R.r0 = c->r0; R.r1 = c->r1;
        unsigned int *sh = reinterpret_cast<unsigned int *>(s_keys);
        for (;;) {
            R.shift = sel_shift(R.klo, R.khi);
            for (int q = threadIdx.x; q < SEL_BINS; q += blockDim.x) sh[q] = 0;
            __syncthreads();
            for (long long i = threadIdx.x; i < P.N; i += blockDim.x) {
                if (!B.alive[i]) continue;
                const unsigned long long key = dkey(X[i]);
                if (key >= R.klo && key <= R.khi) atomicAdd(&sh[(key - R.klo) >> R.shift], 1u);
            }
            __syncthreads();
            const int shift = R.shift;
            sel_scan_narrow(sh, false, R, s_scan, s_res);
            if (shift == 0) break;
        }
        v0 = R.klo; v1 = R.khi;
    }
    if (threadIdx.x == 0) {
        const double a = dunkey(v0), b = dunkey(v1), g = c->gamma;
        double eps;
        if (dfinite(a) && dfinite(b)) eps = xadd(a, xmul(g, xsub(b, a)));
        else eps = xadd(xmul(xsub(1.0, g), a), xmul(g, b));
        c->eps = eps;
        c->xmin = dunkey(c->xmin_key);
        c->flag = (eps > c->xmin) ? 0 : 1; // ref :136-141
        c->cand_count = 0;
        c->tk_final = 0;
        c->sel_done = 0;
    }
}

// ------------------------------------------------------------------ alive cut + ESS + resample decision, ref :136-147
// block b owns particles [1024 b, 1024 b + 1024): 256 threads x 4 consecutive particles
constexpr int CUT_THREADS = 256;
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

__global__ void __launch_bounds__(CUT_THREADS) k_alive_cut(SmcBufs B, SmcParams P, int nblocks) {
    __shared__ unsigned int s_w[33];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    const double *X = B.X[c->cur];
    const double eps = c->eps;
    const int flag = c->flag;
    const long long base = ((long long)blockIdx.x * CUT_THREADS + threadIdx.x) * 4;
    unsigned int cnt = 0;
    if (base + 3 < P.N && (reinterpret_cast<unsigned long long>(X) & 15ull) == 0) {
        const double2 x0 = *reinterpret_cast<const double2 *>(X + base), x1 = *reinterpret_cast<const double2 *>(X + base + 2);
        uchar4 a;
        a.x = flag ? (x0.x <= eps) : (x0.x < eps); a.y = flag ? (x0.y <= eps) : (x0.y < eps);
        a.z = flag ? (x1.x <= eps) : (x1.x < eps); a.w = flag ? (x1.y <= eps) : (x1.y < eps);
        *reinterpret_cast<uchar4 *>(B.alive + base) = a;
        cnt = a.x + a.y + a.z + a.w;
    } else {
        for (int q = 0; q < 4; ++q)
            if (base + q < P.N) {
                const double x = X[base + q];
                const unsigned int a = flag ? (x <= eps) : (x < eps);
                B.alive[base + q] = (unsigned char)a;
                cnt += a;
            }
    }
    unsigned int total;
    block_excl_scan_256(cnt, s_w, total);
    if (threadIdx.x == 0) B.blockcnt[blockIdx.x] = total;
    if (!last_block(&c->tk_cut)) return;
    // last block: exclusive scan of the per-block counts (in place), total = ESS
    unsigned long long carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += CUT_THREADS) {
        const int q = b0 + threadIdx.x;
        const unsigned int v = q < nblocks ? __ldcg(&B.blockcnt[q]) : 0u;
        unsigned int chunk;
        const unsigned int excl = block_excl_scan_256(v, s_w, chunk);
        if (q < nblocks) B.blockcnt[q] = (unsigned int)(carry + excl);
        carry += chunk;
    }
    if (threadIdx.x == 0) {
        const long long ess = (long long)carry;
        c->ess = ess;
        c->tk_cut = 0;
        c->bounds_known = 1; // from now on every alive cost is <= eps
        // ref :145  alpha*ESS <= nparticles*min_r_ess, FP64, exactly these operands
        c->resample = xmul(P.alpha, (double)ess) <= xmul((double)P.N, P.min_r_ess);
        if (c->resample && ess == 0) c->err = KABC_ERR_DEGENERATE;
        c->n_alive = c->resample ? P.N : ess; // ref :151-152: after resampling everything is alive
        if (c->resample) c->resampled_log = 1;
    }
}

// idxalive = (1:N)[alive], ref :146; `alive .= true` (ref :152) is applied here as well
__global__ void __launch_bounds__(CUT_THREADS) k_resample_scatter(SmcBufs B, SmcParams P) {
    __shared__ unsigned int s_w[33];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || !c->resample) return;
    const long long base = ((long long)blockIdx.x * CUT_THREADS + threadIdx.x) * 4;
    unsigned int a[4] = {0u, 0u, 0u, 0u};
    if (base + 3 < P.N) {
        const uchar4 av = *reinterpret_cast<const uchar4 *>(B.alive + base);
        a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
        *reinterpret_cast<uchar4 *>(B.alive + base) = make_uchar4(1, 1, 1, 1);
    } else {
        for (int q = 0; q < 4; ++q)
            if (base + q < P.N) { a[q] = B.alive[base + q]; B.alive[base + q] = 1; }
    }
    unsigned int total;
    unsigned int pos = B.blockcnt[blockIdx.x] + block_excl_scan_256(a[0] + a[1] + a[2] + a[3], s_w, total);
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (a[q]) B.idxalive[pos++] = (unsigned int)(base + q);
}

// One row of the population, pushed into every peer replica of copy `dst` (multi GPU, peer memory over NVLink).
// Every row of a copy has exactly ONE writer in the whole job -- its owner rank -- so the pushes need no ordering
// against anything but the end-of-sweep barrier.
template <int DM>
__device__ __forceinline__ void push_row(const SmcBufs &B, const SmcParams &P, int dst, long long i, const double (&th_row)[DM],
                                         double X, double lp) {
    const long long N = P.N;
    const long long base = (long long)dst * (P.d + 2) * N;
    for (int r = 0; r < B.n_peers; ++r) {
        if (r == P.rank) continue;
        double *q = B.peer[r] + base;
        q[(long long)P.d * N + i] = X;
        if (B.shard_rows) continue;
#pragma unroll
        for (int k = 0; k < DM; ++k)
            if (k < P.d) q[(long long)k * N + i] = th_row[k];
        q[(long long)(P.d + 1) * N + i] = lp;
    }
}

// ------------------------------------------------------------------ resample gather + propose, ref :147-152, :160-167, :172-175
// Reads the complete copy S = cur (rows through the resampling map idx[k] = idxalive[k mod n], ref :146-147, or the
// identity when the reference does not resample) and writes copy D = cur^1: every owned particle's row is
// materialised in D here (the physical gather of the reference, done by the owner only); particles that go on to the
// simulator are appended to the work list and finalised by the sweep kernel, the others are final here and are
// pushed to the peers.  Partner rows are read from S through the same map, so no rank ever needs another rank's D rows.
template <int DM> // DM >= d: compile-time bound of the parameter loops (rows stay in registers)
__global__ void __launch_bounds__(256, 5) // 48 registers: the kernel is latency-bound on its gathers, occupancy matters
k_smc_propose(SmcBufs B, SmcParams P, DPriors pri, RoundKeys rk, long long lo, long long hi, double sqrt_np) {
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c) || c->retry_done) return;
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = P.N;
    const int S = c->cur, D = S ^ 1;
    const double *th = B.th[S];
    const int resample = c->resample;
    const unsigned int n_src = (unsigned int)c->ess;
    const uint32_t epoch = c->epoch;
    bool push = false, defer = false;
    int dec = 0;
    long long a = -1, b = -1;
    double z = dnan(), lprob = dnan(), lpip = dnan();
    if (i < hi) {
        // the particle's own row after the (possible) resampling
        const long long ri = resample ? (long long)B.idxalive[(unsigned int)i % n_src] : i;
        const bool alive_i = resample ? true : (B.alive[i] != 0);
        // where row r of copy S lives: the local replica, or (sharded rows) the slab of the rank that owns r
        const long long per = N / P.world, sbase = (long long)S * (P.d + 2) * N;
        auto src_of = [&](long long r) -> const double * {
            return B.shard_rows ? B.peer[(int)(r / per)] + sbase : th;
        };
        double row[DM];
        const double *own = src_of(ri);
#pragma unroll
        for (int k = 0; k < DM; ++k) {
            row[k] = 0.0;
            if (k < P.d) {
                row[k] = own[(long long)k * N + ri];
                B.th[D][(long long)k * N + i] = row[k];
            }
        }
        const double Xi = B.X[S][ri], lpi_i = own[(long long)(P.d + 1) * N + ri];
        B.X[D][i] = Xi;
        B.lpi[D][i] = lpi_i;
        if (alive_i) {
            Stream st(rk, ST_PROPOSE, (uint32_t)i, epoch);
            a = i; b = i;
            while (a == i) a = (long long)index_of(st.next(), (uint32_t)N);
            while (b == i || b == a) b = (long long)index_of(st.next(), (uint32_t)N);
            const long long ra = resample ? (long long)B.idxalive[(unsigned int)a % n_src] : a;
            const long long rb = resample ? (long long)B.idxalive[(unsigned int)b % n_src] : b;
            z = next_normal(st);
            const double sc = xdiv(xmul(P.max_stretch, z), sqrt_np);
            const double *pa = src_of(ra), *pb = src_of(rb);
#pragma unroll
            for (int k = 0; k < DM; ++k) {
                if (k < P.d)
                    B.thp[(long long)k * N + i] = xadd(row[k], xmul(xsub(pb[(long long)k * N + rb], pa[(long long)k * N + ra]), sc));
            }
            const uint32_t wu = st.next();
            const double *thp = B.thp;
            lpip = prior_logpdf_pushed(pri, [&](int k) { return thp[(long long)k * N + i]; });
            if (lpip < 0.0 && !dfinite(lpip)) dec = 1;
            else {
                // ref :174-175  lM = min(lpip - lpi + logcorr, 0); proceed iff log(rand) < lM.
                // log(u) < 0 always (u < 1), so the logarithm is only evaluated when lM < 0 (or when tracing).
                const double lM = fmin(xadd(xsub(lpip, lpi_i), 0.0), 0.0);
                bool pass = true;
                if (!(lM >= 0.0) || B.trace_on) {
                    lprob = xlog(u01(wu));
                    pass = lprob < lM;
                }
                if (!pass) dec = 2;
                else { push = true; B.lpip[i] = lpip; }
            }
            if (B.trace_on && lprob != lprob) lprob = xlog(u01(wu));
        }
        // multi GPU: a row that is final here still has to reach the peers
        if (!push && B.n_peers > 0) {
            if (B.packed) defer = true;
            else push_row<DM>(B, P, D, i, row, Xi, lpi_i);
        }
        if (B.trace_on) {
            B.tr.a[i] = a; B.tr.b[i] = b; B.tr.z[i] = z; B.tr.lprob[i] = lprob; B.tr.lpip[i] = lpip;
            B.tr.dec[i] = (unsigned char)dec; B.tr.xp[i] = dnan();
            if (!alive_i) for (int k = 0; k < P.d; ++k) B.thp[(long long)k * N + i] = dnan();
        }
    }
    // block-aggregated append to the work list: ONE global atomic per CTA (a per-warp atomic on the single counter
    // serialises 32768 requests in L2 and was the top stall of this kernel)
    __shared__ unsigned int s_cnt[8], s_base;
    const unsigned int ball = __ballot_sync(0xffffffffu, push);
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
    if (push) B.work[s_base + s_cnt[warp] + __popc(ball & ((1u << lane) - 1u))] = (unsigned int)i;
    if (B.packed) { // record j of this rank's stream into every peer's inbox: consecutive lanes -> consecutive slots
        __syncthreads();
        const unsigned int ball2 = __ballot_sync(0xffffffffu, defer);
        if (lane == 0) s_cnt[warp] = __popc(ball2);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { const unsigned int v = s_cnt[w]; s_cnt[w] = tot; tot += v; }
            s_base = tot ? atomicAdd(&c->push_count, tot) : 0u;
        }
        __syncthreads();
        if (defer) {
            const long long j = (long long)(s_base + s_cnt[warp] + __popc(ball2 & ((1u << lane) - 1u)));
            const long long per = N / P.world;
            const long long box = B.inbox_off + ((long long)(c->epoch & 1u) * P.world + P.rank) * per * (P.d + 3);
            const double Xi = B.X[D][i], lpi_i = B.lpi[D][i]; // the row was written above (L1/L2 hit)
            for (int r = 0; r < B.n_peers; ++r) {
                if (r == P.rank) continue;
                double *q = B.peer[r] + box;
                q[j] = __longlong_as_double(i);
                for (int k = 0; k < P.d; ++k) q[(long long)(1 + k) * per + j] = B.th[D][(long long)k * N + i];
                q[(long long)(1 + P.d) * per + j] = Xi;
                q[(long long)(2 + P.d) * per + j] = lpi_i;
            }
        }
    }
}

// receiver side of the packed pushes: after the sweep barrier every rank scatters the records its peers left in its
// inbox into copy D (local stores).  blockIdx.y = source rank.
__global__ void __launch_bounds__(256) k_apply_inbox(SmcBufs B, SmcParams P) {
    const SmcCtrl *c = B.ctrl;
    if ((c->honor_stop && c->stop) || c->err || c->retry_done) return;
    const int r = blockIdx.y;
    if (r == P.rank) return;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (long long)B.partial[r].pushed) return;
    const long long N = P.N, per = N / P.world;
    const int D = c->cur ^ 1;
    const double *q = B.peer[P.rank] + B.inbox_off + ((long long)(c->epoch & 1u) * P.world + r) * per * (P.d + 3);
    const long long i = __double_as_longlong(q[j]);
    for (int k = 0; k < P.d; ++k) B.th[D][(long long)k * N + i] = q[(long long)(1 + k) * per + j];
    B.X[D][i] = q[(long long)(1 + P.d) * per + j];
    B.lpi[D][i] = q[(long long)(2 + P.d) * per + j];
}

// ------------------------------------------------------------------ sweep / iteration bookkeeping
// ref :192 (`accepted >= mcmc_tol*nparticles && break`): folds the sweep's counters (this rank's, or every rank's
// after the all-gather) into the control block
__device__ void post_sweep(SmcBufs &B, const SmcParams &P, bool from_partials) {
    SmcCtrl *c = B.ctrl;
    unsigned long long acc = c->sw_accepted, work = c->work_count, ev = c->sw_events, mk = c->sw_minkey;
    if (from_partials) {
        acc = 0; work = 0; ev = 0; mk = ~0ull;
        for (int r = 0; r < P.world; ++r) {
            acc += B.partial[r].accepted; work += B.partial[r].work; ev += B.partial[r].events;
            mk = B.partial[r].minkey < mk ? B.partial[r].minkey : mk;
        }
    }
    c->accepted += acc;
    c->cost_evals += work;
    c->events += ev;
    if (mk < c->xmin_key) c->xmin_key = mk;
    c->sw_accepted = 0; c->sw_events = 0; c->sw_minkey = ~0ull; c->work_count = 0; c->lv_head = 0; c->push_count = 0;
    c->sweeps += 1;
    c->epoch += 1;
    c->cur ^= 1;      // copy D is now complete on every rank
    c->resample = 0;  // a retry sweep of the same iteration reads D as it is (identity map)
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
// what the last block of a sweep kernel does.  mode bit0: fold + close the sweep here (single GPU);
// bit1: also close the iteration (no retries pending); bit2: publish this rank's partials (multi GPU)
__device__ __forceinline__ void sweep_epilogue(SmcBufs &B, const SmcParams &P, int mode) {
    if (threadIdx.x != 0) return;
    SmcCtrl *c = B.ctrl;
    c->tk_sim = 0;
    if (mode & 4) {
        RankPartial p;
        p.accepted = c->sw_accepted; p.work = c->work_count; p.events = c->sw_events; p.minkey = c->sw_minkey;
        p.pushed = c->push_count;
        B.partial[P.rank] = p;
    }
    if (mode & 1) post_sweep(B, P, false);
    if (mode & 2) post_iter(B, P);
}
__global__ void k_post_sweep_dist(SmcBufs B, SmcParams P, int close_iter) {
    if (B.ctrl->honor_stop && B.ctrl->stop) return;
    if (!(B.ctrl->err || B.ctrl->retry_done)) post_sweep(B, P, true);
    if (close_iter) post_iter(B, P);
}
__global__ void k_post_iter(SmcBufs B, SmcParams P) { post_iter(B, P); }
__global__ void k_set_honor_stop(SmcBufs B, int v) { B.ctrl->honor_stop = v; }

// ------------------------------------------------------------------ simulate + accept, ref :176-189
template <int DM>
__device__ __forceinline__ void smc_accept(SmcBufs &B, const SmcParams &P, SmcCtrl *c, long long i, double Xp,
                                           unsigned int &acc) {
    const long long N = P.N;
    const int D = c->cur ^ 1; // the copy this sweep writes (the row was pre-filled by k_smc_propose)
    const bool reject = c->flag ? (Xp > c->eps) : (Xp >= c->eps);
    double row[DM], Xf = 0.0, lpf = 0.0;
#pragma unroll
    for (int k = 0; k < DM; ++k) row[k] = 0.0;
    if (!reject) {
        lpf = B.lpip[i];
        Xf = Xp;
#pragma unroll
        for (int k = 0; k < DM; ++k)
            if (k < P.d) {
                row[k] = B.thp[(long long)k * N + i];
                B.th[D][(long long)k * N + i] = row[k];
            }
        B.X[D][i] = Xf;
        B.lpi[D][i] = lpf;
        acc = 1;
    } else if (B.n_peers > 0) {
#pragma unroll
        for (int k = 0; k < DM; ++k)
            if (k < P.d) row[k] = B.th[D][(long long)k * N + i];
        Xf = B.X[D][i];
        lpf = B.lpi[D][i];
    }
    // multi GPU: the final row (moved or not) goes straight into every peer's replica: NVLink stores that overlap
    // with the other warps' simulation.  Kernel completion flushes them; the NCCL all-gather of the 32-byte partials
    // that closes the sweep is the barrier after which the replicas are read again.
    if (B.n_peers > 0) {
        push_row<DM>(B, P, D, i, row, Xf, lpf);
    }
    if (B.trace_on) { B.tr.xp[i] = Xp; B.tr.dec[i] = reject ? 3 : 4; }
}

template <int KIND, int PREC>
__global__ void __launch_bounds__(256) k_smc_simulate(SmcBufs B, SmcParams P, DModel m, RoundKeys rk, int mode) {
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    if (c->retry_done) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && (mode & 2)) post_iter(B, P);
        return;
    }
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nwork = c->work_count;
    constexpr int DM = KIND == KABC_MODEL_LV_SSA ? 3 : (KIND == KABC_MODEL_DETERMINISTIC ? KABC_MAX_DIM : 2);
    unsigned int nacc_w = 0;
    unsigned long long e_w = 0, key_w = ~0ull;
    if ((w & ~31u) < nwork) {
        unsigned int acc = 0;
        long long ev = 0;
        unsigned long long key = ~0ull;
        if (w < nwork) {
            const long long i = B.work[w];
            const long long N = P.N;
            const double *thp = B.thp;
            double Xp = cost_thread<KIND, PREC>(m, rk, ST_COST, (uint32_t)i, c->epoch, [&](int k) { return thp[(long long)k * N + i]; }, ev);
            smc_accept<DM>(B, P, c, i, Xp, acc);
            if (acc) key = dkey(Xp);
        }
        nacc_w = __popc(__ballot_sync(0xffffffffu, acc));
        e_w = (KIND == KABC_MODEL_LV_SSA) ? warp_sum_u64((unsigned long long)ev) : 0ull;
        if (nacc_w) key_w = warp_min_u64(key);
    }
    // block-level fold of the sweep counters: one set of global atomics per CTA
    __shared__ unsigned int s_acc;
    __shared__ unsigned long long s_ev, s_key;
    if (threadIdx.x == 0) { s_acc = 0; s_ev = 0; s_key = ~0ull; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        if (nacc_w) { atomicAdd(&s_acc, nacc_w); atomicMin(&s_key, key_w); }
        if (e_w) atomicAdd(&s_ev, e_w);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_acc) { atomicAdd(&c->sw_accepted, (unsigned long long)s_acc); atomicMin(&c->sw_minkey, s_key); }
        if (s_ev) atomicAdd(&c->sw_events, s_ev);
    }
    if (last_block(&c->tk_sim)) sweep_epilogue(B, P, mode);
}

// Lotka-Volterra sweep: persistent lanes.  Event counts per trajectory differ by orders of magnitude, so a lane whose
// trajectory ended immediately pulls the next work item (warp-aggregated atomic on the work-list head) instead of
// idling until the slowest lane of its warp finishes.  Per-particle arithmetic is unchanged (LvSim), so results do
// not depend on the schedule.
constexpr int LV_CHUNK = 32; // events between refill checks
template <int PREC>
__global__ void __launch_bounds__(256) k_smc_simulate_lv(SmcBufs B, SmcParams P, DModel m, RoundKeys rk, int mode) {
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    if (c->retry_done) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && (mode & 2)) post_iter(B, P);
        return;
    }
    const unsigned int nwork = c->work_count;
    const long long N = P.N;
    const uint32_t epoch = c->epoch;
    const unsigned int lane = threadIdx.x & 31;
    LvSim<PREC != KABC_F64> sim;
    long long i = -1;
    bool have = false, exhausted = false;
    unsigned int nacc = 0;
    unsigned long long events = 0, key = ~0ull;
    for (;;) {
        const unsigned int need = __ballot_sync(0xffffffffu, !have && !exhausted);
        if (need) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&c->lv_head, (unsigned int)__popc(need));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (!have && !exhausted) {
                const unsigned int w = base + __popc(need & ((1u << lane) - 1u));
                if (w < nwork) {
                    i = B.work[w];
                    sim.init(m, ST_COST, (uint32_t)i, epoch, pushk(m, 0, B.thp[i]), pushk(m, 1, B.thp[N + i]), pushk(m, 2, B.thp[2 * N + i]));
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
            unsigned int acc = 0;
            const double Xp = sim.result();
            smc_accept<3>(B, P, c, i, Xp, acc);
            if (acc) { nacc += 1; const unsigned long long k2 = dkey(Xp); key = k2 < key ? k2 : key; }
            events += (unsigned long long)sim.ev;
            have = false;
        }
    }
    nacc = (unsigned int)warp_sum_u64(nacc);
    events = warp_sum_u64(events);
    key = warp_min_u64(key);
    if (lane == 0) {
        if (nacc) { atomicAdd(&c->sw_accepted, (unsigned long long)nacc); atomicMin(&c->sw_minkey, key); }
        if (events) atomicAdd(&c->sw_events, events);
    }
    if (last_block(&c->tk_sim)) sweep_epilogue(B, P, mode);
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS) k_smc_simulate_gk(SmcBufs B, SmcParams P, DModel m, RoundKeys rk, int mode) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    SmcCtrl *c = B.ctrl;
    if (smc_skip(c)) return;
    if (c->retry_done) {
        if (blockIdx.x == 0 && threadIdx.x == 0 && (mode & 2)) post_iter(B, P);
        return;
    }
    const unsigned int nwork = c->work_count;
    const long long N = P.N;
    for (unsigned int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const long long i = B.work[w];
        const double *thp = B.thp;
        double Xp = cost_gk_block<PREC>(m, rk, ST_COST, (uint32_t)i, c->epoch, pushk(m, 0, thp[i]), pushk(m, 1, thp[N + i]),
                                     pushk(m, 2, thp[2 * N + i]), pushk(m, 3, thp[3 * N + i]), gk_smem);
        if (threadIdx.x == 0) {
            unsigned int acc = 0;
            smc_accept<4>(B, P, c, i, Xp, acc);
            if (acc) { atomicAdd(&c->sw_accepted, 1ull); atomicMin(&c->sw_minkey, dkey(Xp)); }
        }
    }
    if (last_block(&c->tk_sim)) sweep_epilogue(B, P, mode);
}

// recount after kabc_smc_set_state: alive count and the running minimum
__global__ void k_recount(SmcBufs B, long long N) {
    __shared__ unsigned long long s_n, s_min;
    if (threadIdx.x == 0) { s_n = 0; s_min = ~0ull; }
    __syncthreads();
    const double *X = B.X[B.ctrl->cur];
    unsigned long long n = 0, mk = ~0ull;
    for (long long i = threadIdx.x; i < N; i += blockDim.x)
        if (B.alive[i]) {
            n += 1;
            const unsigned long long k = dkey(X[i]);
            mk = k < mk ? k : mk;
        }
    n = warp_sum_u64(n);
    mk = warp_min_u64(mk);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_n, n); atomicMin(&s_min, mk); }
    __syncthreads();
    if (threadIdx.x == 0) {
        B.ctrl->n_alive = (long long)s_n; B.ctrl->ess = (long long)s_n;
        B.ctrl->xmin_key = s_min; B.ctrl->bounds_known = 0;
    }
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
    DevBuf<double> slab, thp, lpip; // slab = both copies of [th | X | lpi]
    std::vector<void *> peer_maps;  // cudaIpcOpenMemHandle mappings to close
    bool p2p = false;               // accepted rows are pushed into the peers' slabs (else: NCCL all-gather)
    DevBuf<unsigned char> alive;
    DevBuf<unsigned int> work, idxalive, blockcnt, hist;
    DevBuf<unsigned long long> cand;
    DevBuf<SmcCtrl> ctrl;
    DevBuf<RankPartial> partial;
    DevBuf<kabc_smc_log_t> log;
    // trace
    DevBuf<long long> ta, tb;
    DevBuf<double> tz, tlprob, tlpip, txp;
    DevBuf<unsigned char> tdec;
    SmcCtrl *h_ctrl = nullptr; // pinned
    long long lo = 0, hi = 0;  // owned particle range [lo,hi) of this rank
    int cur = 0;               // host mirror of ctrl->cur (flips once per iteration)
    bool inited = false;
    long long launches = 0;
    int nblocks_scan = 0;
    // one whole iteration (cut + sweep, all control flow on the device) captured as a CUDA graph: a single launch instead
    // of 7-9, so the short selection kernels run back to back even when the host is not ahead of the device
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

template <int KIND>
static void smc_launch_init_t(kabc_smc *s) {
    const long long n = s->hi - s->lo;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (s->model.precision == KABC_F64)
        k_smc_init<KIND, KABC_F64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->pri, s->model, s->ctx->rk, s->lo, s->hi);
    else
        k_smc_init<KIND, KABC_F32_ACC64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->pri, s->model, s->ctx->rk, s->lo, s->hi);
    SMC_LAUNCHED(s, 1);
}

template <int KIND>
static void smc_launch_sim_t(kabc_smc *s, int mode) {
    const long long n = s->hi - s->lo;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (s->model.precision == KABC_F64)
        k_smc_simulate<KIND, KABC_F64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->model, s->ctx->rk, mode);
    else
        k_smc_simulate<KIND, KABC_F32_ACC64><<<blocks, 256, 0, s->ctx->stream>>>(s->B, s->P, s->model, s->ctx->rk, mode);
    SMC_LAUNCHED(s, 1);
}

static int smc_gk_grid(kabc_smc *s, size_t &smem) {
    smem = gk_smem_bytes(s->model.n_draws, s->model.precision);
    long long cap = (long long)s->ctx->sm_count * gk_blocks_per_sm(s->model.n_draws, s->model.precision);
    long long n = s->hi - s->lo;
    return (int)(n < cap ? n : cap);
}

// all-gather of the per-rank partials and -- unless the rows were already pushed through peer memory -- of the
// ranks' shards of state copy `cur` (theta planes, X, lpi)
static int smc_allgather_state(kabc_smc *s, int cur, bool rows) {
    kabc_ctx *ctx = s->ctx;
    const long long N = s->P.N, per = N / ctx->world;
    if (int rc = nccl_group_start()) return rc;
    if (rows) {
        for (int k = 0; k < s->P.d; ++k)
            if (int rc = nccl_allgather_inplace(ctx, s->B.th[cur] + (long long)k * N, (size_t)per * 8)) return rc;
        if (int rc = nccl_allgather_inplace(ctx, s->B.X[cur], (size_t)per * 8)) return rc;
        if (int rc = nccl_allgather_inplace(ctx, s->B.lpi[cur], (size_t)per * 8)) return rc;
    }
    if (int rc = nccl_allgather_inplace(ctx, s->B.partial, sizeof(RankPartial))) return rc;
    if (int rc = nccl_group_end()) return rc;
    return KABC_OK;
}

// Map every peer's state slab into this process (cudaIpc; NVLink P2P).  The 64-byte handles travel through the
// NCCL communicator the context already owns, so the host language needs no extra plumbing.  On any failure the
// handle stays in the all-gather mode (s->p2p = false): same results, more traffic.
static int smc_attach_peers(kabc_smc *s) {
    kabc_ctx *ctx = s->ctx;
    const int world = ctx->world;
    s->p2p = false;
    s->B.n_peers = 0;
    if (world == 1 || world > KABC_MAX_PEERS) return KABC_OK;
    const char *env = getenv("KABC_NO_P2P");
    const bool want = !(env && env[0] == '1');
    std::vector<cudaIpcMemHandle_t> handles(world);
    memset(handles.data(), 0, sizeof(cudaIpcMemHandle_t) * world);
    int ok = want ? 1 : 0;
    if (ok && cudaIpcGetMemHandle(&handles[ctx->rank], s->slab.p) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    DevBuf<unsigned char> dh;
    DevBuf<unsigned long long> dok;
    KABC_CUDA_TRY(dh.alloc(sizeof(cudaIpcMemHandle_t) * world));
    KABC_CUDA_TRY(dok.alloc(1));
    KABC_CUDA_TRY(cudaMemcpyAsync(dh.p + sizeof(cudaIpcMemHandle_t) * ctx->rank, &handles[ctx->rank], sizeof(cudaIpcMemHandle_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = nccl_allgather_inplace(ctx, dh.p, sizeof(cudaIpcMemHandle_t))) return rc;
    KABC_CUDA_TRY(cudaMemcpyAsync(handles.data(), dh.p, sizeof(cudaIpcMemHandle_t) * world, cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<void *> maps(world, nullptr);
    for (int r = 0; ok && r < world; ++r) {
        if (r == ctx->rank) { maps[r] = s->slab.p; continue; }
        if (cudaIpcOpenMemHandle(&maps[r], handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // every rank must take the same path: agree through a sum
    unsigned long long okv = (unsigned long long)ok;
    KABC_CUDA_TRY(cudaMemcpyAsync(dok.p, &okv, 8, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = nccl_allreduce_sum_u64(ctx, dok.p, 1)) return rc;
    KABC_CUDA_TRY(cudaMemcpyAsync(&okv, dok.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if ((int)okv == world) {
        for (int r = 0; r < world; ++r) {
            s->B.peer[r] = (double *)maps[r];
            if (r != ctx->rank) s->peer_maps.push_back(maps[r]);
        }
        s->B.n_peers = world;
        s->p2p = true;
        // measured on 4 x B200: fine-grained remote partner reads cost more than replicating the rows
        // (propose 311 us vs 165 us at 2^22 particles), so full-row replication is the default
        const char *e2 = getenv("KABC_SHARD_ROWS");
        s->B.shard_rows = (e2 && e2[0] == '1') ? 1 : 0;
        // measured on B200s (2^20 particles per GPU): at 2 GPUs the direct row stores are as fast (0.925 vs 0.926 ms per
        // iteration), at 4 GPUs the packed records win (propose 128 us vs 165 us, 0.98 vs 1.01 ms per iteration)
        const char *e3 = getenv("KABC_PACKED_PUSH");
        const bool want_packed = e3 ? (e3[0] == '1') : (world >= 4);
        s->B.packed = (!s->B.shard_rows && want_packed) ? 1 : 0;
        s->B.inbox_off = 2 * (long long)(s->P.d + 2) * s->P.N;
    } else {
        for (int r = 0; r < world; ++r)
            if (r != ctx->rank && maps[r]) cudaIpcCloseMemHandle(maps[r]);
    }
    return KABC_OK;
}

static int smc_enqueue_init(kabc_smc *s) {
    kabc_ctx *ctx = s->ctx;
    k_smc_reset<<<4, 1024, 0, ctx->stream>>>(s->B, s->P);
    SMC_LAUNCHED(s, 1);
    s->cur = 0;
    switch (s->model.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: smc_launch_init_t<KABC_MODEL_NORMAL_MEANSTD>(s); break;
    case KABC_MODEL_MA2_AUTOCOV: smc_launch_init_t<KABC_MODEL_MA2_AUTOCOV>(s); break;
    case KABC_MODEL_LV_SSA: smc_launch_init_t<KABC_MODEL_LV_SSA>(s); break;
    case KABC_MODEL_DETERMINISTIC: smc_launch_init_t<KABC_MODEL_DETERMINISTIC>(s); break;
    case KABC_MODEL_SOCKS: smc_launch_init_t<KABC_MODEL_SOCKS>(s); break;
    case KABC_MODEL_GK_OCTILE: {
        const long long n = s->hi - s->lo;
        k_smc_init_prior<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk, s->lo, s->hi);
        size_t smem;
        int grid = smc_gk_grid(s, smem);
        if (s->model.precision == KABC_F64) {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_smc_init_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_smc_init_gk<KABC_F64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->model, ctx->rk, s->lo, s->hi);
        } else {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_smc_init_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_smc_init_gk<KABC_F32_ACC64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->model, ctx->rk, s->lo, s->hi);
        }
        SMC_LAUNCHED(s, 2);
        break;
    }
    }
    KABC_CUDA_TRY(cudaGetLastError());
    if (ctx->world > 1) {
        k_smc_write_partial<<<1, 1, 0, ctx->stream>>>(s->B, s->P);
        if (int rc = smc_allgather_state(s, 0, true)) return rc;
        KABC_CUDA_TRY(cudaMemsetAsync(s->B.alive, 1, (size_t)s->P.N, ctx->stream));
        k_smc_post_init<<<1, 1, 0, ctx->stream>>>(s->B, s->P, 1);
        SMC_LAUNCHED(s, 2);
    } else {
        k_smc_post_init<<<1, 1, 0, ctx->stream>>>(s->B, s->P, 0);
        SMC_LAUNCHED(s, 1);
    }
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

// one MCMC sweep: propose -> (work list) -> simulate+accept (+ bookkeeping in the last block)
static int smc_enqueue_sweep(kabc_smc *s, bool close_iter) {
    kabc_ctx *ctx = s->ctx;
    const long long n = s->hi - s->lo;
    const bool dist = ctx->world > 1;
    const int mode = dist ? 4 : (1 | (close_iter ? 2 : 0));
    {
        const unsigned pb = (unsigned)((n + 255) / 256);
        const double sq = sqrt((double)s->P.d);
        if (s->P.d <= 2) k_smc_propose<2><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk, s->lo, s->hi, sq);
        else if (s->P.d <= 4) k_smc_propose<4><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk, s->lo, s->hi, sq);
        else k_smc_propose<KABC_MAX_DIM><<<pb, 256, 0, ctx->stream>>>(s->B, s->P, s->pri, ctx->rk, s->lo, s->hi, sq);
    }
    SMC_LAUNCHED(s, 1);
    s->mark();
    switch (s->model.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: smc_launch_sim_t<KABC_MODEL_NORMAL_MEANSTD>(s, mode); break;
    case KABC_MODEL_MA2_AUTOCOV: smc_launch_sim_t<KABC_MODEL_MA2_AUTOCOV>(s, mode); break;
    case KABC_MODEL_LV_SSA: {
        long long nb = (n + 255) / 256, cap = (long long)ctx->sm_count * 8;
        const unsigned blocks = (unsigned)(nb < cap ? nb : cap);
        if (s->model.precision == KABC_F64)
            k_smc_simulate_lv<KABC_F64><<<blocks, 256, 0, ctx->stream>>>(s->B, s->P, s->model, ctx->rk, mode);
        else
            k_smc_simulate_lv<KABC_F32_ACC64><<<blocks, 256, 0, ctx->stream>>>(s->B, s->P, s->model, ctx->rk, mode);
        SMC_LAUNCHED(s, 1);
        break;
    }
    case KABC_MODEL_DETERMINISTIC: smc_launch_sim_t<KABC_MODEL_DETERMINISTIC>(s, mode); break;
    case KABC_MODEL_SOCKS: smc_launch_sim_t<KABC_MODEL_SOCKS>(s, mode); break;
    case KABC_MODEL_GK_OCTILE: {
        size_t smem;
        int grid = smc_gk_grid(s, smem);
        if (s->model.precision == KABC_F64) {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_smc_simulate_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_smc_simulate_gk<KABC_F64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->model, ctx->rk, mode);
        } else {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_smc_simulate_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_smc_simulate_gk<KABC_F32_ACC64><<<grid, GK_THREADS, smem, ctx->stream>>>(s->B, s->P, s->model, ctx->rk, mode);
        }
        SMC_LAUNCHED(s, 1);
        break;
    }
    }
    s->mark();
    if (dist) {
        // barrier + counters; without peer memory also the rows this rank wrote into copy D
        if (int rc = smc_allgather_state(s, s->cur ^ 1, !s->p2p)) return rc;
        if (s->B.packed) {
            const dim3 grid((unsigned)((n + 255) / 256), (unsigned)ctx->world);
            k_apply_inbox<<<grid, 256, 0, ctx->stream>>>(s->B, s->P);
            SMC_LAUNCHED(s, 1);
        }
        k_post_sweep_dist<<<1, 1, 0, ctx->stream>>>(s->B, s->P, close_iter ? 1 : 0);
        SMC_LAUNCHED(s, 1);
        s->mark();
    }
    s->cur ^= 1; // host mirror of ctrl->cur: every executed sweep writes the other copy
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

static int smc_enqueue_cut(kabc_smc *s) {
    kabc_ctx *ctx = s->ctx;
    const long long N = s->P.N;
    int sel_blocks = (int)((N + SEL_THREADS * 16 - 1) / (SEL_THREADS * 16));
    if (sel_blocks > ctx->sm_count * 2) sel_blocks = ctx->sm_count * 2;
    if (sel_blocks < 1) sel_blocks = 1;
    s->mark();
    k_sel_hist<0><<<sel_blocks, SEL_THREADS, 0, ctx->stream>>>(s->B, s->P);
    s->mark();
    k_sel_hist<1><<<sel_blocks, SEL_THREADS, 0, ctx->stream>>>(s->B, s->P);
    s->mark();
    k_sel_final<<<sel_blocks, SEL_THREADS, 0, ctx->stream>>>(s->B, s->P);
    s->mark();
    k_alive_cut<<<s->nblocks_scan, CUT_THREADS, 0, ctx->stream>>>(s->B, s->P, s->nblocks_scan);
    s->mark();
    k_resample_scatter<<<s->nblocks_scan, CUT_THREADS, 0, ctx->stream>>>(s->B, s->P);
    s->mark();
    SMC_LAUNCHED(s, 5);
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
            // ref :192 -- leave the retry loop as soon as enough moves were accepted
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
// host between sweeps; no NCCL row all-gather, whose buffers alternate), else kernel by kernel
static int smc_launch_iteration(kabc_smc *s) {
    kabc_ctx *ctx = s->ctx;
    static const bool env_off = [] { const char *e = getenv("KABC_NO_GRAPH"); return e && e[0] == '1'; }();
    const bool eligible = s->graph_ok && !env_off && !s->prof && s->P.mcmc_retrys == 0 &&
                          s->model.kind != KABC_MODEL_GK_OCTILE && (ctx->world == 1 || s->p2p);
    if (!eligible) return smc_enqueue_iteration(s);
    if (!s->iter_graph) {
        const long long l0 = s->launches, c0 = ctx->launches;
        const int cur0 = s->cur;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        int rc = KABC_OK;
        if (e == cudaSuccess) {
            rc = smc_enqueue_iteration(s);
            e = cudaStreamEndCapture(ctx->stream, &g);
        }
        s->graph_kernels = (int)(s->launches - l0);
        s->launches = l0; ctx->launches = c0; s->cur = cur0; // nothing ran yet
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
    s->cur ^= 1;
    return KABC_OK;
}

extern "C" {

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
    s->ctx = ctx; s->pri = pri; s->model = m; s->cfg = *cfg;
    s->P.N = N; s->P.d = d; s->P.alpha = cfg->alpha; s->P.mcmc_tol = cfg->mcmc_tol; s->P.epstol = cfg->epstol;
    s->P.r_epstol = cfg->r_epstol; s->P.min_r_ess = cfg->min_r_ess; s->P.max_stretch = cfg->max_stretch;
    s->P.mcmc_retrys = cfg->mcmc_retrys; s->P.max_iterations = cfg->max_iterations;
    s->P.rank = ctx->rank; s->P.world = ctx->world;
    s->lo = N / ctx->world * ctx->rank;
    s->hi = N / ctx->world * (ctx->rank + 1);
    s->nblocks_scan = (int)((N + SCAN_THREADS - 1) / SCAN_THREADS);
    const size_t nd = (size_t)N * d;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    const size_t copy_elems = nd + 2 * (size_t)N; // [th | X | lpi]
    // multi GPU: + the inbox of the packed pushes, 2 sweep parities x world sources x (N/world) records of d+3 doubles
    if (ctx->world > 1) A(s->slab.alloc(2 * copy_elems + 2 * (size_t)N * (d + 3))); // peer-mapped slabs are not recycled
    else A(s->slab.alloc(ctx, 2 * copy_elems));
    A(s->thp.alloc(ctx, nd)); A(s->lpip.alloc(ctx, N));
    A(s->alive.alloc(ctx, N)); A(s->work.alloc(ctx, N)); A(s->idxalive.alloc(ctx, N)); A(s->blockcnt.alloc(ctx, s->nblocks_scan));
    A(s->hist.alloc(ctx, SEL_BINS)); A(s->cand.alloc(ctx, SEL_CAP)); A(s->ctrl.alloc(ctx, 1)); A(s->partial.alloc(ctx, ctx->world));
    const long long log_cap = 1 << 14;
    A(s->log.alloc(ctx, log_cap));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&s->h_ctrl, sizeof(SmcCtrl));
    if (e != cudaSuccess) {
        delete s;
        return set_error(KABC_ERR_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
    }
    memset(s->h_ctrl, 0, sizeof(SmcCtrl));
    for (int c = 0; c < 2; ++c) {
        s->B.th[c] = s->slab.p + (size_t)c * copy_elems;
        s->B.X[c] = s->B.th[c] + nd;
        s->B.lpi[c] = s->B.X[c] + N;
    }
    memset(s->B.peer, 0, sizeof s->B.peer);
    s->B.n_peers = 0;
    s->B.shard_rows = 0;
    s->B.packed = 0;
    s->B.inbox_off = 0;
    s->B.alive = s->alive.p; s->B.thp = s->thp.p;
    s->B.lpip = s->lpip.p; s->B.work = s->work.p; s->B.idxalive = s->idxalive.p; s->B.blockcnt = s->blockcnt.p;
    s->B.hist = s->hist.p; s->B.cand = s->cand.p; s->B.ctrl = s->ctrl.p; s->B.partial = s->partial.p;
    s->B.log = s->log.p; s->B.log_cap = log_cap;
    memset(&s->B.tr, 0, sizeof s->B.tr);
    s->B.trace_on = 0;
    KABC_CUDA_TRY(cudaMemsetAsync(s->ctrl.p, 0, sizeof(SmcCtrl), ctx->stream));
    if (int rc = smc_attach_peers(s)) {
        std::string keep = g_last_error;
        kabc_smc_destroy(s);
        g_last_error = keep;
        return rc;
    }
    *out = s;
    return KABC_OK;
}

int kabc_smc_destroy(kabc_smc_t *s) {
    if (!s) return KABC_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    smc_drop_graph(s);
    for (void *m : s->peer_maps) cudaIpcCloseMemHandle(m);
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
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const int cur = s->cur;
    if (s->B.shard_rows && s->inited) {
        // theta / lpi rows are only valid on their owner: assemble the full state (collective: every rank calls this)
        if (int rc = smc_allgather_state(s, cur, true)) return rc;
    }
    const size_t N = (size_t)s->P.N;
    cudaStream_t st = s->ctx->stream;
    if (theta) KABC_CUDA_TRY(cudaMemcpyAsync(theta, s->B.th[cur], 8 * N * s->P.d, cudaMemcpyDeviceToHost, st));
    if (X) KABC_CUDA_TRY(cudaMemcpyAsync(X, s->B.X[cur], 8 * N, cudaMemcpyDeviceToHost, st));
    if (lpi) KABC_CUDA_TRY(cudaMemcpyAsync(lpi, s->B.lpi[cur], 8 * N, cudaMemcpyDeviceToHost, st));
    if (alive) KABC_CUDA_TRY(cudaMemcpyAsync(alive, s->B.alive, N, cudaMemcpyDeviceToHost, st));
    KABC_CUDA_TRY(cudaStreamSynchronize(st));
    return KABC_OK;
}

int kabc_smc_set_state(kabc_smc_t *s, const double *theta, const double *X, const double *lpi, const uint8_t *alive) {
    if (!s) return set_error(KABC_ERR_INVALID_ARG, "smc is NULL");
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    const int cur = s->cur;
    const size_t N = (size_t)s->P.N;
    cudaStream_t st = s->ctx->stream;
    if (theta) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.th[cur], theta, 8 * N * s->P.d, cudaMemcpyHostToDevice, st));
    if (X) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.X[cur], X, 8 * N, cudaMemcpyHostToDevice, st));
    if (lpi) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.lpi[cur], lpi, 8 * N, cudaMemcpyHostToDevice, st));
    if (alive) KABC_CUDA_TRY(cudaMemcpyAsync(s->B.alive, alive, N, cudaMemcpyHostToDevice, st));
    k_recount<<<1, 1024, 0, st>>>(s->B, s->P.N);
    SMC_LAUNCHED(s, 1);
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
    KABC_CUDA_TRY(cudaSetDevice(s->ctx->device));
    KABC_CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    if (on && !s->ta.p) {
        const size_t N = (size_t)s->P.N;
        KABC_CUDA_TRY(s->ta.alloc(N)); KABC_CUDA_TRY(s->tb.alloc(N)); KABC_CUDA_TRY(s->tz.alloc(N));
        KABC_CUDA_TRY(s->tlprob.alloc(N)); KABC_CUDA_TRY(s->tlpip.alloc(N)); KABC_CUDA_TRY(s->txp.alloc(N));
        KABC_CUDA_TRY(s->tdec.alloc(N));
        s->B.tr.a = s->ta.p; s->B.tr.b = s->tb.p; s->B.tr.z = s->tz.p; s->B.tr.lprob = s->tlprob.p;
        s->B.tr.lpip = s->tlpip.p; s->B.tr.xp = s->txp.p; s->B.tr.dec = s->tdec.p; s->B.tr.thp = s->thp.p;
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
        // device when `stop` was set (smc_skip), and the host then takes its buffer flip back.
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
                s->cur ^= 1;                             // ... so its host-side buffer flip is undone
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
    if (!rc) rc = kabc_smc_get_state(s, out_theta, out_cost, nullptr, out_alive);
    if (!rc && out_theta) { // ref src/smc.jl:200: the returned particles are push_p(prior, .)
        const long long N = s->P.N;
        for (int k = 0; k < d; ++k)
            if (prior_is_discrete(s->pri.p[k]))
                for (long long i = 0; i < N; ++i) out_theta[(long long)k * N + i] = nearbyint(out_theta[(long long)k * N + i]);
    }
    if (!rc) {
        if (out_eps) *out_eps = s->h_ctrl->eps;
        if (out_iterations) *out_iterations = s->h_ctrl->iteration;
        if (out_cost_evals) *out_cost_evals = (int64_t)s->h_ctrl->cost_evals;
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
