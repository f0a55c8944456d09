// kabc_core.cu -- context, descriptor ingestion, prior kernels, bare cost evaluation, pipe microbenchmarks.
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>
#include "kabc_host.hpp"
#include "kabc_gk.cuh"
#include "kabc_nccl.hpp"

namespace kabc {

thread_local std::string g_last_error;

int set_error(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// NVTX3 is header-only: the calls dispatch to the profiler's injection library when one is attached and are no-ops otherwise
void nvtx_push(const char *name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

static double std_normal_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

int ingest_priors(const kabc_prior_t *prior, int d, DPriors &out) {
    if (!prior) return set_error(KABC_ERR_INVALID_ARG, "prior is NULL");
    if (d < 1 || d > KABC_MAX_DIM) return set_error(KABC_ERR_INVALID_ARG, "d must be in 1..%d", KABC_MAX_DIM);
    out.d = d;
    for (int k = 0; k < d; ++k) {
        const kabc_prior_t &p = prior[k];
        DPrior &q = out.p[k];
        q.kind = p.kind;
        q.p0 = p.p0; q.p1 = p.p1; q.lo = p.lo; q.hi = p.hi;
        q.c0 = 0.0; q.c1 = 0.0; q.c2 = 0.0;
        switch (p.kind) {
        case KABC_PRIOR_UNIFORM:
            if (!(p.p1 > p.p0)) return set_error(KABC_ERR_INVALID_ARG, "Uniform(a,b) needs a < b (component %d)", k);
            q.c0 = log(p.p1 - p.p0);
            break;
        case KABC_PRIOR_NORMAL:
            if (!(p.p1 > 0)) return set_error(KABC_ERR_INVALID_ARG, "Normal(mu,sigma) needs sigma > 0 (component %d)", k);
            q.c0 = log(p.p1);
            break;
        case KABC_PRIOR_TRUNC_NORMAL:
            if (!(p.p1 > 0)) return set_error(KABC_ERR_INVALID_ARG, "Normal(mu,sigma) needs sigma > 0 (component %d)", k);
            if (!(p.hi > p.lo)) return set_error(KABC_ERR_INVALID_ARG, "Truncated(...,lo,hi) needs lo < hi (component %d)", k);
            q.c0 = log(p.p1);
            q.c1 = log(std_normal_cdf((p.hi - p.p0) / p.p1) - std_normal_cdf((p.lo - p.p0) / p.p1));
            break;
        case KABC_PRIOR_BETA:
            if (!(p.p0 > 0 && p.p1 > 0) || !std::isfinite(p.p0) || !std::isfinite(p.p1))
                return set_error(KABC_ERR_INVALID_ARG, "Beta(a,b) needs a, b > 0 (component %d)", k);
            q.c0 = (lgamma(p.p0) + lgamma(p.p1)) - lgamma(p.p0 + p.p1);
            break;
        case KABC_PRIOR_NEG_BINOMIAL:
            if (!(p.p0 > 0) || !std::isfinite(p.p0) || !(p.p1 > 0 && p.p1 <= 1))
                return set_error(KABC_ERR_INVALID_ARG, "NegativeBinomial(r,p) needs r > 0 and 0 < p <= 1 (component %d)", k);
            q.c0 = lgamma(p.p0);
            q.c1 = p.p0 * log(p.p1);
            q.c2 = log1p(-p.p1);
            break;
        case KABC_PRIOR_DISCRETE_UNIFORM:
            if (p.p0 != floor(p.p0) || p.p1 != floor(p.p1) || !(p.p1 >= p.p0) || !(p.p1 - p.p0 < 2147483647.0))
                return set_error(KABC_ERR_INVALID_ARG, "DiscreteUniform(a,b) needs integers a <= b (component %d)", k);
            q.c0 = log((p.p1 - p.p0) + 1.0);
            break;
        default:
            return set_error(KABC_ERR_INVALID_ARG, "unknown prior kind %d (component %d)", p.kind, k);
        }
    }
    return KABC_OK;
}

int ingest_model(const kabc_model_t *model, int d, DModel &out) {
    if (!model) return set_error(KABC_ERR_INVALID_ARG, "model is NULL");
    out.kind = model->kind;
    out.push_mask = 0;
    out.precision = model->precision;
    out.n_draws = model->n_draws;
    out.n_target = model->n_target;
    memcpy(out.target, model->target, sizeof out.target);
    memcpy(out.param, model->param, sizeof out.param);
    if (model->precision != KABC_F64 && model->precision != KABC_F32_ACC64)
        return set_error(KABC_ERR_INVALID_ARG, "unknown precision %d", model->precision);
    switch (model->kind) {
    case KABC_MODEL_NORMAL_MEANSTD:
        if (d != 2) return set_error(KABC_ERR_INVALID_ARG, "normal model needs d = 2 (mu, sigma)");
        if (model->n_draws < 2) return set_error(KABC_ERR_INVALID_ARG, "normal model needs n_draws >= 2");
        break;
    case KABC_MODEL_MA2_AUTOCOV:
        if (d != 2) return set_error(KABC_ERR_INVALID_ARG, "MA(2) model needs d = 2");
        if (model->n_draws < 3) return set_error(KABC_ERR_INVALID_ARG, "MA(2) model needs n_draws >= 3");
        break;
    case KABC_MODEL_GK_OCTILE:
        if (d != 4) return set_error(KABC_ERR_INVALID_ARG, "g-and-k model needs d = 4 (A, B, g, k)");
        if (model->n_draws < 8 || model->n_draws > KABC_GK_MAX_DRAWS)
            return set_error(KABC_ERR_INVALID_ARG, "g-and-k model needs 8 <= n_draws <= %d", KABC_GK_MAX_DRAWS);
        break;
    case KABC_MODEL_LV_SSA:
        if (d != 3) return set_error(KABC_ERR_INVALID_ARG, "Lotka-Volterra model needs d = 3 (log rates)");
        if (!(model->param[3] >= 1 && model->param[3] <= KABC_MAX_TARGET / 2))
            return set_error(KABC_ERR_INVALID_ARG, "Lotka-Volterra grid size param[3] must be in 1..%d", KABC_MAX_TARGET / 2);
        break;
    case KABC_MODEL_DETERMINISTIC:
        if (model->param[0] == 2.0 && d != 2)
            return set_error(KABC_ERR_INVALID_ARG, "deterministic model with param0 = 2 needs d = 2 (n, du)");
        break;
    case KABC_MODEL_SOCKS:
        if (d != 2) return set_error(KABC_ERR_INVALID_ARG, "socks model needs d = 2 (n_socks, prop_pairs)");
        if (!(model->param[0] >= 1 && model->param[0] <= KABC_SOCKS_MAX_PICKED) || model->param[0] != floor(model->param[0]))
            return set_error(KABC_ERR_INVALID_ARG, "socks model needs 1 <= n_picked (param0) <= %d", KABC_SOCKS_MAX_PICKED);
        if (model->n_target != 2) return set_error(KABC_ERR_INVALID_ARG, "socks model needs n_target = 2 (pairs, odds)");
        break;
    default:
        return set_error(KABC_ERR_INVALID_ARG, "unknown model kind %d", model->kind);
    }
    return KABC_OK;
}

// ------------------------------------------------------------------ kernels
__global__ void k_prior_logpdf(DPriors P, const double *__restrict__ th, long long n, double *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = prior_logpdf(P, [&](int k) { return th[(long long)k * n + i]; });
}

__global__ void k_prior_sample(DPriors P, RoundKeys rk, long long n, uint32_t first_id, uint32_t epoch,
                               double *__restrict__ th, int *__restrict__ bad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Stream st(rk, ST_PRIOR, first_id + (uint32_t)i, epoch);
    bool ok = true;
    for (int k = 0; k < P.d; ++k) {
        double x;
        ok &= prior1_sample(P.p[k], st, x);
        th[(long long)k * n + i] = x;
    }
    if (!ok) atomicExch(bad, 1);
}

template <int KIND, int PREC>
__global__ void __launch_bounds__(256)
k_eval_cost(DModel m, RoundKeys rk, const double *__restrict__ th, long long n, uint32_t first_id, uint32_t epoch,
            double *__restrict__ out, long long *__restrict__ out_events) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long ev;
    double c = cost_thread<KIND, PREC>(m, rk, ST_COST, first_id + (uint32_t)i, epoch,
                                       [&](int k) { return th[(long long)k * n + i]; }, ev);
    out[i] = c;
    if (out_events) out_events[i] = ev;
}

template <int PREC>
__global__ void __launch_bounds__(GK_THREADS, (PREC == KABC_F64 ? 1 : 3))
k_eval_cost_gk(DModel m, RoundKeys rk, const double *__restrict__ th, long long n, uint32_t first_id, uint32_t epoch,
               double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    for (long long i = blockIdx.x; i < n; i += gridDim.x) {
        double c = cost_gk_block<PREC>(m, rk, ST_COST, first_id + (uint32_t)i, epoch, th[i], th[n + i], th[2 * n + i],
                                       th[3 * n + i], gk_smem);
        if (threadIdx.x == 0) out[i] = c;
    }
}

// list-driven variant (ABCDE / pfilter): entry w of `list` names a particle i of an N-particle SoA state; its cost is
// drawn from stream (tag, i, epoch) and written to out[i].  The entry count lives in device memory.
template <int KIND, int PREC>
__global__ void __launch_bounds__(256)
k_eval_cost_list(DModel m, RoundKeys rk, const double *__restrict__ th, long long N, const unsigned int *__restrict__ list,
                 const unsigned int *__restrict__ count, uint32_t tag, uint32_t epoch, double *__restrict__ out) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= *count) return;
    const long long i = list[w];
    long long ev;
    out[i] = cost_thread<KIND, PREC>(m, rk, tag, (uint32_t)i, epoch, [&](int k) { return th[(long long)k * N + i]; }, ev);
}
template <int PREC>
__global__ void __launch_bounds__(GK_THREADS, (PREC == KABC_F64 ? 1 : 3))
k_eval_cost_list_gk(DModel m, RoundKeys rk, const double *__restrict__ th, long long N, const unsigned int *__restrict__ list,
                    const unsigned int *__restrict__ count, uint32_t tag, uint32_t epoch, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char gk_smem[];
    const unsigned int n = *count;
    for (unsigned int w = blockIdx.x; w < n; w += gridDim.x) {
        const long long i = list[w];
        double c = cost_gk_block<PREC>(m, rk, tag, (uint32_t)i, epoch, th[i], th[N + i], th[2 * N + i], th[3 * N + i], gk_smem);
        if (threadIdx.x == 0) out[i] = c;
    }
}

template <int KIND>
static void launch_eval_list(kabc_ctx *ctx, const DModel &m, const double *d_th, long long N, const unsigned int *list,
                             const unsigned int *count, long long max_count, uint32_t tag, uint32_t epoch, double *d_out) {
    const unsigned blocks = (unsigned)((max_count + 255) / 256);
    if (m.precision == KABC_F64)
        k_eval_cost_list<KIND, KABC_F64><<<blocks, 256, 0, ctx->stream>>>(m, ctx->rk, d_th, N, list, count, tag, epoch, d_out);
    else
        k_eval_cost_list<KIND, KABC_F32_ACC64><<<blocks, 256, 0, ctx->stream>>>(m, ctx->rk, d_th, N, list, count, tag, epoch, d_out);
    ctx->launches += 1;
}

int eval_cost_list_device(kabc_ctx *ctx, const DModel &m, const double *d_th, long long N, const unsigned int *list,
                          const unsigned int *count, long long max_count, uint32_t tag, uint32_t epoch, double *d_out) {
    if (max_count <= 0) return KABC_OK;
    switch (m.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: launch_eval_list<KABC_MODEL_NORMAL_MEANSTD>(ctx, m, d_th, N, list, count, max_count, tag, epoch, d_out); break;
    case KABC_MODEL_MA2_AUTOCOV: launch_eval_list<KABC_MODEL_MA2_AUTOCOV>(ctx, m, d_th, N, list, count, max_count, tag, epoch, d_out); break;
    case KABC_MODEL_LV_SSA: launch_eval_list<KABC_MODEL_LV_SSA>(ctx, m, d_th, N, list, count, max_count, tag, epoch, d_out); break;
    case KABC_MODEL_DETERMINISTIC: launch_eval_list<KABC_MODEL_DETERMINISTIC>(ctx, m, d_th, N, list, count, max_count, tag, epoch, d_out); break;
    case KABC_MODEL_SOCKS: launch_eval_list<KABC_MODEL_SOCKS>(ctx, m, d_th, N, list, count, max_count, tag, epoch, d_out); break;
    case KABC_MODEL_GK_OCTILE: {
        size_t smem = gk_smem_bytes(m.n_draws, m.precision);
        long long cap = (long long)ctx->sm_count * gk_blocks_per_sm(m.n_draws, m.precision);
        unsigned blocks = (unsigned)(max_count < cap ? max_count : cap);
        if (m.precision == KABC_F64) {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_eval_cost_list_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_eval_cost_list_gk<KABC_F64><<<blocks, GK_THREADS, smem, ctx->stream>>>(m, ctx->rk, d_th, N, list, count, tag, epoch, d_out);
        } else {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_eval_cost_list_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_eval_cost_list_gk<KABC_F32_ACC64><<<blocks, GK_THREADS, smem, ctx->stream>>>(m, ctx->rk, d_th, N, list, count, tag, epoch, d_out);
        }
        ctx->launches += 1;
        break;
    }
    default: return set_error(KABC_ERR_INVALID_ARG, "unknown model kind %d", m.kind);
    }
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

template <int KIND>
static int launch_eval(kabc_ctx *ctx, const DModel &m, const double *d_th, long long n, uint32_t first_id,
                       uint32_t epoch, double *d_out, long long *d_ev) {
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    if (m.precision == KABC_F64)
        k_eval_cost<KIND, KABC_F64><<<blocks, threads, 0, ctx->stream>>>(m, ctx->rk, d_th, n, first_id, epoch, d_out, d_ev);
    else
        k_eval_cost<KIND, KABC_F32_ACC64><<<blocks, threads, 0, ctx->stream>>>(m, ctx->rk, d_th, n, first_id, epoch, d_out, d_ev);
    ctx->launches += 1;
    return KABC_OK;
}

int eval_cost_device(kabc_ctx *ctx, const DModel &m, const double *d_th, long long n, uint32_t first_id, uint32_t epoch,
                     double *d_out, long long *d_ev) {
    if (n == 0) return KABC_OK;
    switch (m.kind) {
    case KABC_MODEL_NORMAL_MEANSTD: launch_eval<KABC_MODEL_NORMAL_MEANSTD>(ctx, m, d_th, n, first_id, epoch, d_out, d_ev); break;
    case KABC_MODEL_MA2_AUTOCOV: launch_eval<KABC_MODEL_MA2_AUTOCOV>(ctx, m, d_th, n, first_id, epoch, d_out, d_ev); break;
    case KABC_MODEL_LV_SSA: launch_eval<KABC_MODEL_LV_SSA>(ctx, m, d_th, n, first_id, epoch, d_out, d_ev); break;
    case KABC_MODEL_DETERMINISTIC: launch_eval<KABC_MODEL_DETERMINISTIC>(ctx, m, d_th, n, first_id, epoch, d_out, d_ev); break;
    case KABC_MODEL_SOCKS: launch_eval<KABC_MODEL_SOCKS>(ctx, m, d_th, n, first_id, epoch, d_out, d_ev); break;
    case KABC_MODEL_GK_OCTILE: {
        size_t smem = gk_smem_bytes(m.n_draws, m.precision);
        long long cap = (long long)ctx->sm_count * gk_blocks_per_sm(m.n_draws, m.precision);
        unsigned blocks = (unsigned)(n < cap ? n : cap);
        if (m.precision == KABC_F64) {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_eval_cost_gk<KABC_F64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_eval_cost_gk<KABC_F64><<<blocks, GK_THREADS, smem, ctx->stream>>>(m, ctx->rk, d_th, n, first_id, epoch, d_out);
        } else {
            KABC_CUDA_TRY(cudaFuncSetAttribute(k_eval_cost_gk<KABC_F32_ACC64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_eval_cost_gk<KABC_F32_ACC64><<<blocks, GK_THREADS, smem, ctx->stream>>>(m, ctx->rk, d_th, n, first_id, epoch, d_out);
        }
        ctx->launches += 1;
        if (d_ev) KABC_CUDA_TRY(cudaMemsetAsync(d_ev, 0, sizeof(long long) * (size_t)n, ctx->stream));
        break;
    }
    default: return set_error(KABC_ERR_INVALID_ARG, "unknown model kind %d", m.kind);
    }
    KABC_CUDA_TRY(cudaGetLastError());
    return KABC_OK;
}

// ------------------------------------------------------------------ pipe microbenchmarks
// Each thread runs ITER iterations of UNROLL independent dependent-chains; rate = threads*ITER*UNROLL*ops / time.
template <int KIND>
__global__ void __launch_bounds__(256) k_microbench(RoundKeys rk, int iters, float *sink_f, double *sink_d) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (KIND == 0) { // FFMA
        float a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = (float)(tid + q) * 1e-9f;
        const float m1 = 0.999f, c = 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = __fmaf_rn(a[q], m1, c);
        }
        float s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += a[q];
        if (s == 12345.678f) sink_f[0] = s;
    } else if (KIND == 1 || KIND == 3) { // IMAD (lo) / LOP3
        uint32_t a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = tid * 2654435761u + q;
        uint32_t m1 = tid | 1u, c = tid ^ 0x5bd1e995u;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (KIND == 1) a[q] = a[q] * m1 + c;
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[q]) : "r"(m1), "r"(c));
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s ^= a[q];
        if (s == 0x12345u) sink_f[0] = (float)s;
    } else if (KIND == 2) { // IMAD.WIDE: 32x32 -> 64, both halves consumed (Philox pattern)
        uint32_t a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = tid * 2654435761u + q;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                unsigned long long p = (unsigned long long)0xD2511F53u * a[q];
                a[q] = (uint32_t)(p >> 32) + (uint32_t)p; // +1 IADD per op (reported separately)
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s ^= a[q];
        if (s == 0x12345u) sink_f[0] = (float)s;
    } else if (KIND >= 4 && KIND <= 6) { // MUFU
        float a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = 1.0f + (float)((tid + q) & 1023) * 1e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (KIND == 4) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[q]));
                if (KIND == 5) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[q]));
                if (KIND == 6) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[q]));
            }
        }
        float s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += a[q];
        if (s == 12345.678f) sink_f[0] = s;
    } else if (KIND == 7) { // DFMA
        double a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = (double)(tid + q) * 1e-9;
        const double m1 = 0.999, c = 1e-3;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = __fma_rn(a[q], m1, c);
        }
        double s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += a[q];
        if (s == 12345.678) sink_d[0] = s;
    } else if (KIND == 8) { // I2F.U32 (+ F2I back so the chain stays integer: reported as pairs)
        uint32_t a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = tid * 2654435761u + q;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float f;
                asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a[q]));
                a[q] = __float_as_uint(f) ^ a[q];
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s ^= a[q];
        if (s == 0x12345u) sink_f[0] = (float)s;
    } else if (KIND == 9) { // Philox words
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t w0, w1, w2, w3;
            philox4x32_10(rk, (uint32_t)it, tid, 0u, 9u, w0, w1, w2, w3);
            acc ^= w0 ^ w1 ^ w2 ^ w3;
        }
        if (acc == 0x12345u) sink_f[0] = (float)acc;
    } else if (KIND == 10) { // FP32 Box-Muller normals incl. Philox
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t w0, w1, w2, w3;
            philox4x32_10(rk, (uint32_t)it, tid, 0u, 10u, w0, w1, w2, w3);
            float z0, z1, z2, z3;
            normal_pair32(w0, w1, z0, z1);
            normal_pair32(w2, w3, z2, z3);
            s0 += z0; s1 += z1; s2 += z2; s3 += z3;
        }
        float s = (s0 + s1) + (s2 + s3);
        if (s == 12345.678f) sink_f[0] = s;
    } else if (KIND == 11) { // FP64 spec normals incl. Philox
        double s0 = 0, s1 = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t w0, w1, w2, w3;
            philox4x32_10(rk, (uint32_t)it, tid, 0u, 11u, w0, w1, w2, w3);
            double z0, z1, z2, z3;
            normal_pair64(w0, w1, z0, z1);
            normal_pair64(w2, w3, z2, z3);
            s0 += z0 + z1; s1 += z2 + z3;
        }
        double s = s0 + s1;
        if (s == 12345.678) sink_d[0] = s;
    } else if (KIND == 12) { // Philox + the four u32->f32 conversions only (isolates I2FP)
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t w0, w1, w2, w3;
            philox4x32_10(rk, (uint32_t)it, tid, 0u, 12u, w0, w1, w2, w3);
            s0 += __uint2float_rn(w0); s1 += __uint2float_rn(w1); s2 += __uint2float_rn(w2); s3 += __uint2float_rn(w3);
        }
        float s = (s0 + s1) + (s2 + s3);
        if (s == 12345.678f) sink_f[0] = s;
    } else if (KIND == 13) { // Box-Muller with mantissa-trick conversions (no I2FP): 23-bit uniforms
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t w[4];
            philox4x32_10(rk, (uint32_t)it, tid, 0u, 13u, w[0], w[1], w[2], w[3]);
            float z[4];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float f1 = __uint_as_float(0x3f800000u | (w[2 * q] >> 9));     // [1,2)
                float f2 = __uint_as_float(0x3f800000u | (w[2 * q + 1] >> 9)); // [1,2)
                float u1 = __fsub_rn(2.0f, f1);                                // (0,1]
                float ang = __fmaf_rn(f2, 6.2831853071795865f, -6.2831853071795865f);
                float r = mufu_sqrt(__fmul_rn(mufu_lg2(u1), -1.3862943611198906f));
                z[2 * q] = __fmul_rn(r, mufu_cos(ang));
                z[2 * q + 1] = __fmul_rn(r, mufu_sin(ang));
            }
            s0 += z[0]; s1 += z[1]; s2 += z[2]; s3 += z[3];
        }
        float s = (s0 + s1) + (s2 + s3);
        if (s == 12345.678f) sink_f[0] = s;
    } else if (KIND == 14) { // Box-Muller, I2FP kept, no accumulation FADD chain differences: MUFU count halved (lg2+sqrt only)
        float s0 = 0, s1 = 0;
        for (int it = 0; it < iters; ++it) {
            uint32_t w0, w1, w2, w3;
            philox4x32_10(rk, (uint32_t)it, tid, 0u, 14u, w0, w1, w2, w3);
            float u1 = __fmaf_rn(__uint2float_rn(w0), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
            float u2 = __fmaf_rn(__uint2float_rn(w2), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
            s0 += mufu_sqrt(__fmul_rn(mufu_lg2(u1), -1.3862943611198906f)) * __uint_as_float(0x3f800000u | (w1 >> 9));
            s1 += mufu_sqrt(__fmul_rn(mufu_lg2(u2), -1.3862943611198906f)) * __uint_as_float(0x3f800000u | (w3 >> 9));
        }
        float s = s0 + s1;
        if (s == 12345.678f) sink_f[0] = s;
    }
}

} // namespace kabc

using namespace kabc;

// =================================================================== C ABI
extern "C" {

int kabc_version(void) { return KABC_VERSION; }
const char *kabc_last_error(void) { return g_last_error.c_str(); }

int kabc_device_count(int *count) {
    if (!count) return set_error(KABC_ERR_INVALID_ARG, "count is NULL");
    KABC_CUDA_TRY(cudaGetDeviceCount(count));
    return KABC_OK;
}

int kabc_nccl_unique_id(char id[KABC_NCCL_ID_BYTES]) { return nccl_unique_id(id); }

static int ctx_create_common(int device, uint64_t seed, int rank, int world, kabc_ctx **out) {
    if (!out) return set_error(KABC_ERR_INVALID_ARG, "ctx out pointer is NULL");
    if (world < 1 || rank < 0 || rank >= world) return set_error(KABC_ERR_INVALID_ARG, "bad rank/world %d/%d", rank, world);
    if (world > KABC_MAX_PEERS) return set_error(KABC_ERR_INVALID_ARG, "at most %d ranks per job", KABC_MAX_PEERS);
    int ndev = 0;
    KABC_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return set_error(KABC_ERR_CUDA, "no CUDA device: libkissabc_cuda has no CPU fallback");
    if (device < 0 || device >= ndev) return set_error(KABC_ERR_INVALID_ARG, "device %d out of range (%d devices)", device, ndev);
    KABC_CUDA_TRY(cudaSetDevice(device));
    kabc_ctx *ctx = new kabc_ctx();
    ctx->device = device;
    ctx->seed = seed;
    ctx->rk = make_round_keys(seed);
    ctx->rank = rank;
    ctx->world = world;
    cudaDeviceProp prop;
    KABC_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    KABC_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    KABC_CUDA_TRY(cudaEventCreate(&ctx->ev0));
    KABC_CUDA_TRY(cudaEventCreate(&ctx->ev1));
    KABC_CUDA_TRY(cudaMalloc((void **)&ctx->xseq, sizeof(unsigned long long)));
    KABC_CUDA_TRY(cudaMemset(ctx->xseq, 0, sizeof(unsigned long long)));
    ctx->arena_attached = (world == 1);
    *out = ctx;
    return KABC_OK;
}

int kabc_ctx_create_dist(int device, uint64_t seed, int rank, int world, const char id[KABC_NCCL_ID_BYTES], kabc_ctx_t **out) {
    kabc_ctx *ctx = nullptr;
    if (int rc = ctx_create_common(device, seed, rank, world, &ctx)) return rc;
    if (world > 1) {
        int rc = nccl_comm_init(ctx, id);
        if (rc) { kabc_ctx_destroy(ctx); return rc; }
    }
    *out = ctx;
    return KABC_OK;
}

int kabc_ctx_create(int device, uint64_t seed, kabc_ctx_t **out) {
    return kabc_ctx_create_dist(device, seed, 0, 1, nullptr, out);
}

int kabc_ctx_create_ranks(int device, uint64_t seed, int rank, int world, kabc_ctx_t **out) {
    kabc_ctx *ctx = nullptr;
    if (int rc = ctx_create_common(device, seed, rank, world, &ctx)) return rc;
    ctx->host_exchange = true;
    *out = ctx;
    return KABC_OK;
}

} // extern "C"

namespace kabc {
// ------------------------------------------------------------------ peer arena
static void arena_unmap(kabc_ctx *ctx) {
    for (int r = 0; r < KABC_MAX_PEERS; ++r) {
        if (r != ctx->rank && ctx->arena_map[r]) cudaIpcCloseMemHandle(ctx->arena_map[r]);
        ctx->arena_map[r] = nullptr;
    }
    ctx->arena_attached = (ctx->world == 1);
}

static int arena_allocate(kabc_ctx *ctx, size_t total_bytes) {
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    arena_unmap(ctx);
    if (ctx->arena) {
        // peers unmap before the owner frees (cudaIpc contract): a host-side barrier sits between the two
        if (ctx->comm) {
            DevBuf<unsigned long long> one;
            KABC_CUDA_TRY(one.alloc(1));
            KABC_CUDA_TRY(cudaMemsetAsync(one.p, 0, 8, ctx->stream));
            if (int rc = nccl_allreduce_sum_u64(ctx, one.p, 1)) return rc;
            KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        }
        cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
    }
    total_bytes = (total_bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    cudaError_t e = cudaMalloc((void **)&ctx->arena, total_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ctx->trim(0);
        e = cudaMalloc((void **)&ctx->arena, total_bytes);
    }
    if (e != cudaSuccess) { ctx->arena = nullptr; return set_error(KABC_ERR_CUDA, "peer arena allocation of %zu bytes failed: %s", total_bytes, cudaGetErrorString(e)); }
    ctx->arena_bytes = total_bytes;
    ctx->arena_top = KABC_ARENA_HEADER;
    ctx->arena_users = 0;
    KABC_CUDA_TRY(cudaMemset(ctx->arena, 0, KABC_ARENA_HEADER));
    KABC_CUDA_TRY(cudaMemset(ctx->xseq, 0, sizeof(unsigned long long)));
    KABC_CUDA_TRY(cudaDeviceSynchronize());
    ctx->arena_map[ctx->rank] = ctx->arena;
    return KABC_OK;
}

static int arena_open(kabc_ctx *ctx, const cudaIpcMemHandle_t *handles) {
    for (int r = 0; r < ctx->world; ++r) {
        if (r == ctx->rank) { ctx->arena_map[r] = ctx->arena; continue; }
        void *m = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&m, handles[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            arena_unmap(ctx);
            return set_error(KABC_ERR_CUDA, "cudaIpcOpenMemHandle of rank %d's arena failed: %s (peer access over NVLink is required)", r,
                             cudaGetErrorString(e));
        }
        ctx->arena_map[r] = m;
    }
    ctx->arena_attached = true;
    return KABC_OK;
}

int arena_reserve(kabc_ctx *ctx, size_t bytes) {
    const size_t need = KABC_ARENA_HEADER + ((bytes + 255) & ~(size_t)255) + 256;
    if (ctx->world == 1) return set_error(KABC_ERR_STATE, "single-rank contexts have no peer arena");
    if (ctx->arena && ctx->arena_attached && ctx->arena_bytes >= need) return KABC_OK;
    if (ctx->host_exchange)
        return set_error(KABC_ERR_STATE, ctx->arena_attached ? "peer arena too small: need %zu bytes, have %zu (pass a larger size to kabc_ctx_arena_export)"
                                                             : "peer arena not attached: call kabc_ctx_arena_export / kabc_ctx_arena_attach first (need %zu bytes, have %zu)",
                         need, ctx->arena_bytes);
    if (ctx->arena_users > 0) return set_error(KABC_ERR_STATE, "peer arena too small and still in use by another handle");
    if (int rc = arena_allocate(ctx, need)) return rc;
    // move the 64-byte cudaIpc handles through the NCCL communicator the context already owns
    const int world = ctx->world;
    std::vector<cudaIpcMemHandle_t> handles(world);
    memset(handles.data(), 0, sizeof(cudaIpcMemHandle_t) * world);
    KABC_CUDA_TRY(cudaIpcGetMemHandle(&handles[ctx->rank], ctx->arena));
    DevBuf<unsigned char> dh;
    KABC_CUDA_TRY(dh.alloc(sizeof(cudaIpcMemHandle_t) * world));
    KABC_CUDA_TRY(cudaMemcpyAsync(dh.p + sizeof(cudaIpcMemHandle_t) * ctx->rank, &handles[ctx->rank], sizeof(cudaIpcMemHandle_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = nccl_allgather_inplace(ctx, dh.p, sizeof(cudaIpcMemHandle_t))) return rc;
    KABC_CUDA_TRY(cudaMemcpyAsync(handles.data(), dh.p, sizeof(cudaIpcMemHandle_t) * world, cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return arena_open(ctx, handles.data());
}

int arena_alloc(kabc_ctx *ctx, size_t bytes, size_t *offset) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (ctx->arena_users == 0) {
        ctx->arena_top = KABC_ARENA_HEADER;
        if (int rc = arena_reserve(ctx, bytes)) return rc;
    } else if (!ctx->arena || ctx->arena_top + bytes > ctx->arena_bytes) {
        return set_error(KABC_ERR_STATE, "peer arena exhausted: destroy the other multi-rank handle first");
    }
    *offset = ctx->arena_top;
    ctx->arena_top += bytes;
    ctx->arena_users += 1;
    return KABC_OK;
}

void arena_release(kabc_ctx *ctx) {
    if (ctx->arena_users > 0) ctx->arena_users -= 1;
    if (ctx->arena_users == 0) ctx->arena_top = KABC_ARENA_HEADER;
}

} // namespace kabc

extern "C" {

int kabc_ctx_arena_export(kabc_ctx_t *ctx, uint64_t bytes, char handle[KABC_IPC_HANDLE_BYTES]) {
    if (!ctx || !handle) return set_error(KABC_ERR_INVALID_ARG, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == KABC_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    if (ctx->world == 1) return set_error(KABC_ERR_STATE, "single-rank contexts have no peer arena");
    if (ctx->arena_users > 0) return set_error(KABC_ERR_STATE, "the peer arena is in use");
    if (int rc = arena_allocate(ctx, KABC_ARENA_HEADER + (size_t)bytes + 512)) return rc;
    cudaIpcMemHandle_t h;
    KABC_CUDA_TRY(cudaIpcGetMemHandle(&h, ctx->arena));
    memcpy(handle, &h, sizeof h);
    return KABC_OK;
}

int kabc_ctx_arena_attach(kabc_ctx_t *ctx, const char *handles) {
    if (!ctx || !handles) return set_error(KABC_ERR_INVALID_ARG, "NULL argument");
    if (!ctx->arena) return set_error(KABC_ERR_STATE, "kabc_ctx_arena_export must be called first");
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<cudaIpcMemHandle_t> hs(ctx->world);
    memcpy(hs.data(), handles, sizeof(cudaIpcMemHandle_t) * ctx->world);
    return arena_open(ctx, hs.data());
}

int kabc_ctx_destroy(kabc_ctx_t *ctx) {
    if (!ctx) return KABC_OK;
    cudaSetDevice(ctx->device);
    if (ctx->in_use_count() > 0 || ctx->arena_users > 0)
        return set_error(KABC_ERR_STATE, "destroy the smc / ais handles of a context before the context");
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    arena_unmap(ctx);
    if (ctx->comm) nccl_comm_destroy(ctx);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->xseq) cudaFree(ctx->xseq);
    for (auto &e : ctx->cache) cudaFree(e.p);
    ctx->cache.clear();
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return KABC_OK;
}

int kabc_ctx_info(const kabc_ctx_t *ctx, int *device, int *rank, int *world, int *sm_count) {
    if (!ctx) return set_error(KABC_ERR_INVALID_ARG, "ctx is NULL");
    if (device) *device = ctx->device;
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    if (sm_count) *sm_count = ctx->sm_count;
    return KABC_OK;
}

int kabc_prior_logpdf(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, const double *theta, int64_t n, double *out) {
    if (!ctx || !theta || !out || n < 0) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    DPriors P;
    if (int rc = ingest_priors(prior, d, P)) return rc;
    if (n == 0) return KABC_OK;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf<double> dth, dout;
    KABC_CUDA_TRY(dth.alloc((size_t)n * d));
    KABC_CUDA_TRY(dout.alloc((size_t)n));
    KABC_CUDA_TRY(cudaMemcpyAsync(dth.p, theta, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, ctx->stream));
    k_prior_logpdf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(P, dth.p, n, dout.p);
    ctx->launches += 1;
    KABC_CUDA_TRY(cudaGetLastError());
    KABC_CUDA_TRY(cudaMemcpyAsync(out, dout.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return KABC_OK;
}

int kabc_prior_sample(kabc_ctx_t *ctx, const kabc_prior_t *prior, int d, int64_t n, uint32_t first_id, uint32_t epoch,
                      double *out_theta) {
    if (!ctx || !out_theta || n < 0) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    DPriors P;
    if (int rc = ingest_priors(prior, d, P)) return rc;
    if (n == 0) return KABC_OK;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf<double> dth;
    DevBuf<int> dbad;
    KABC_CUDA_TRY(dth.alloc((size_t)n * d));
    KABC_CUDA_TRY(dbad.alloc(1));
    KABC_CUDA_TRY(cudaMemsetAsync(dbad.p, 0, sizeof(int), ctx->stream));
    k_prior_sample<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(P, ctx->rk, n, first_id, epoch, dth.p, dbad.p);
    ctx->launches += 1;
    KABC_CUDA_TRY(cudaGetLastError());
    int bad = 0;
    KABC_CUDA_TRY(cudaMemcpyAsync(out_theta, dth.p, sizeof(double) * (size_t)n * d, cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaMemcpyAsync(&bad, dbad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (bad) return set_error(KABC_ERR_INVALID_ARG, "prior sampling failed (truncation too extreme)");
    return KABC_OK;
}

int kabc_eval_cost_device(kabc_ctx_t *ctx, const kabc_model_t *model, int d, const double *d_theta, int64_t n,
                          uint32_t first_id, uint32_t epoch, double *d_out, float *out_ms) {
    if (!ctx || !d_theta || !d_out || n < 0) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    DModel m;
    if (int rc = ingest_model(model, d, m)) return rc;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    KABC_CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    if (int rc = eval_cost_device(ctx, m, d_theta, n, first_id, epoch, d_out, nullptr)) return rc;
    KABC_CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (out_ms) KABC_CUDA_TRY(cudaEventElapsedTime(out_ms, ctx->ev0, ctx->ev1));
    return KABC_OK;
}

int kabc_eval_cost(kabc_ctx_t *ctx, const kabc_model_t *model, int d, const double *theta, int64_t n, uint32_t first_id,
                   uint32_t epoch, double *out_cost, int64_t *out_events) {
    if (!ctx || !theta || !out_cost || n < 0) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    DModel m;
    if (int rc = ingest_model(model, d, m)) return rc;
    if (n == 0) return KABC_OK;
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf<double> dth, dout;
    DevBuf<long long> dev;
    KABC_CUDA_TRY(dth.alloc((size_t)n * d));
    KABC_CUDA_TRY(dout.alloc((size_t)n));
    if (out_events) KABC_CUDA_TRY(dev.alloc((size_t)n));
    KABC_CUDA_TRY(cudaMemcpyAsync(dth.p, theta, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = eval_cost_device(ctx, m, dth.p, n, first_id, epoch, dout.p, dev.p)) return rc;
    KABC_CUDA_TRY(cudaMemcpyAsync(out_cost, dout.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_events)
        KABC_CUDA_TRY(cudaMemcpyAsync(out_events, dev.p, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return KABC_OK;
}

int kabc_microbench(kabc_ctx_t *ctx, int kind, double *out_rate, float *out_ms) {
    if (!ctx || kind < 0 || kind > 14) return set_error(KABC_ERR_INVALID_ARG, "bad argument");
    KABC_CUDA_TRY(cudaSetDevice(ctx->device));
    DevBuf<float> sf;
    DevBuf<double> sd;
    KABC_CUDA_TRY(sf.alloc(1));
    KABC_CUDA_TRY(sd.alloc(1));
    const int threads = 256, blocks = ctx->sm_count * 8;
    int iters = 4096;
    double ops_per_iter = 8.0;
    if (kind == 7) iters = 2048;
    if (kind == 9) { iters = 1024; ops_per_iter = 4.0; }
    if (kind == 10) { iters = 1024; ops_per_iter = 4.0; }
    if (kind == 11) { iters = 64; ops_per_iter = 4.0; }
    if (kind >= 12) { iters = 1024; ops_per_iter = 4.0; }
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        KABC_CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        switch (kind) {
#define KABC_MB(K) case K: k_microbench<K><<<blocks, threads, 0, ctx->stream>>>(ctx->rk, iters, sf.p, sd.p); break;
            KABC_MB(0) KABC_MB(1) KABC_MB(2) KABC_MB(3) KABC_MB(4) KABC_MB(5) KABC_MB(6) KABC_MB(7) KABC_MB(8) KABC_MB(9)
            KABC_MB(10) KABC_MB(11) KABC_MB(12) KABC_MB(13) KABC_MB(14)
#undef KABC_MB
        }
        KABC_CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        KABC_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        KABC_CUDA_TRY(cudaGetLastError());
        float ms = 0;
        KABC_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    ctx->launches += 4;
    if (out_ms) *out_ms = best;
    if (out_rate) *out_rate = (double)blocks * threads * (double)iters * ops_per_iter / (best * 1e-3);
    return KABC_OK;
}

} // extern "C"
