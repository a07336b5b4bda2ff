// K5: per-segment outputs over the final split points.
//
// Replaces LogMarginalLikelyhoodComputer.scores / mean_counts
// (/root/reference/src/pasio/log_marginal_likelyhood.py:67-74, :80-83) and feeds
// NopSplitter.split (/root/reference/src/pasio/splitters/nop_splitter.py:15-18).
//   score_k = (Ga[c_k] - (c_k + alpha) * Lg[len_k]) + segment_creation_cost
//   mean_k  = c_k / len_k                       (int / int true division == one IEEE divide)
#include "common.cuh"

namespace {

template <bool AI>
__global__ void segment_scores_kernel(const int32_t *__restrict__ cand, i64 m, const i64 *__restrict__ cg,
                                      const double *__restrict__ ga, const double *__restrict__ lg,
                                      i64 alpha_int, double alpha, double pen,
                                      double *__restrict__ scores, i64 *__restrict__ segcounts,
                                      double *__restrict__ means)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m - 1; k += (i64)gridDim.x * blockDim.x) {
        const i64 a = cand ? (i64)__ldg(cand + k) : k;
        const i64 b = cand ? (i64)__ldg(cand + k + 1) : k + 1;
        const i64 cnt = __ldg(cg + b) - __ldg(cg + a);
        const i64 len = b - a;
        const double shifted = AI ? __ll2double_rn(cnt + alpha_int) : __dadd_rn(__ll2double_rn(cnt), alpha);
        const double self = __dsub_rn(__ldg(ga + cnt), __dmul_rn(shifted, __ldg(lg + len)));
        if (scores) scores[k] = __dadd_rn(self, pen);
        if (segcounts) segcounts[k] = cnt;
        if (means) means[k] = __ddiv_rn(__ll2double_rn(cnt), __ll2double_rn(len));
    }
}

__global__ void gather_i64_kernel(const i64 *__restrict__ src, const int32_t *__restrict__ idx32,
                                  const i64 *__restrict__ idx64, i64 m, i64 *__restrict__ out)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (i64)gridDim.x * blockDim.x) {
        const i64 p = idx32 ? (i64)idx32[k] : (idx64 ? idx64[k] : k);
        out[k] = __ldg(src + p);
    }
}

__global__ void gather_f64_kernel(const double *__restrict__ src, const int32_t *__restrict__ idx32, i64 m,
                                  double *__restrict__ out)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (i64)gridDim.x * blockDim.x) {
        const i64 p = idx32 ? (i64)idx32[k] : k;
        out[k] = src[p];
    }
}

__global__ void lmm_kernel(const double *__restrict__ scores, const double *__restrict__ logfac_full,
                           const int32_t *__restrict__ cand, i64 m, double *__restrict__ lmm)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m - 1; k += (i64)gridDim.x * blockDim.x) {
        const i64 a = cand ? (i64)__ldg(cand + k) : k;
        const i64 b = cand ? (i64)__ldg(cand + k + 1) : k + 1;
        lmm[k] = __dsub_rn(scores[k], __dsub_rn(logfac_full[b], logfac_full[a]));
    }
}

inline unsigned grid_for(pasio_ctx *ctx, i64 n)
{
    i64 g = (n + 255) / 256;
    i64 cap = (i64)ctx->sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace

// numpy's pairwise summation, leaf level: one thread per leaf of <= 128 elements (the host walks the recursion that
// splits n at multiples of 8 and combines the leaf sums in the same order): 8 running sums, then the fixed tree,
// then the remainder -- operation for operation numpy's DOUBLE_pairwise_sum.
__global__ void pairwise_leaf_kernel(const double *__restrict__ a, const i64 *__restrict__ leaf_start, i64 n_leaves,
                                     double *__restrict__ leaf_sum)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_leaves) return;
    const double *x = a + leaf_start[t];
    const i64 n = leaf_start[t + 1] - leaf_start[t];
    double res;
    if (n < 8) {
        res = 0.;
        for (i64 i = 0; i < n; ++i) res = __dadd_rn(res, x[i]);
    } else {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = x[k];
        i64 i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], x[i + k]);
        }
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __dadd_rn(res, x[i]);
    }
    leaf_sum[t] = res;
}

int launch_pairwise_leaves(pasio_ctx *ctx, const double *d_values, const i64 *d_leaf_start, i64 n_leaves, double *d_leaf_sum)
{
    TimingScope ts(ctx, TF_SCORE);
    pairwise_leaf_kernel<<<(unsigned)((n_leaves + 127) / 128), 128, 0, ctx->stream>>>(d_values, d_leaf_start, n_leaves, d_leaf_sum);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_segment_scores(pasio_ctx *ctx, double *d_scores, i64 *d_segcounts, double *d_means)
{
    const double *ga = ctx->tab[PASIO_TAB_LGAMMA_ALPHA].as<double>();
    const double *lg = ctx->tab[PASIO_TAB_LOG].as<double>();
    TimingScope ts(ctx, TF_SCORE);
    if (ctx->alpha_is_int)
        segment_scores_kernel<true><<<grid_for(ctx, ctx->m), 256, 0, ctx->stream>>>(
            cur_cand(ctx), ctx->m, ctx->cg.as<i64>(), ga, lg, ctx->alpha_int, ctx->alpha, ctx->pen, d_scores,
            d_segcounts, d_means);
    else
        segment_scores_kernel<false><<<grid_for(ctx, ctx->m), 256, 0, ctx->stream>>>(
            cur_cand(ctx), ctx->m, ctx->cg.as<i64>(), ga, lg, ctx->alpha_int, ctx->alpha, ctx->pen, d_scores,
            d_segcounts, d_means);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_gather_i64(pasio_ctx *ctx, const i64 *d_src, const int32_t *d_idx32, const i64 *d_idx64, i64 m, i64 *d_out)
{
    gather_i64_kernel<<<grid_for(ctx, m), 256, 0, ctx->stream>>>(d_src, d_idx32, d_idx64, m, d_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_gather_f64_at_cands(pasio_ctx *ctx, const double *d_src, double *d_out)
{
    gather_f64_kernel<<<grid_for(ctx, ctx->m), 256, 0, ctx->stream>>>(d_src, cur_cand(ctx), ctx->m, d_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_lmm(pasio_ctx *ctx, const double *d_scores, const double *d_logfac_full, double *d_lmm)
{
    TimingScope ts(ctx, TF_SCORE);
    lmm_kernel<<<grid_for(ctx, ctx->m), 256, 0, ctx->stream>>>(d_scores, d_logfac_full, cur_cand(ctx), ctx->m, d_lmm);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}
