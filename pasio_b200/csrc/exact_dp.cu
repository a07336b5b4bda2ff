// K3: the exact O(N^2) SquareSplitter DP for one long candidate list (N up to a few 1e5),
// spread over the whole chip.
//
// Replaces SquareSplitter.split_without_normalizations + collect_split_points
// (/root/reference/src/pasio/splitters/square_splitter.py:67-109) driven by
// all_suffixes_self_score (/root/reference/src/pasio/log_marginal_likelyhood.py:105-132).
//
// Rows are resolved in blocks of XD_ROWS.  For the block starting at row jb:
//   rectangle kernel : columns [0, jb) are final; they are cut into chunks, one CTA per chunk,
//                      lane = row, and each CTA emits a per-row partial (max, first arg-max);
//   diagonal kernel  : one CTA merges the partials in column order (strict '>' keeps the first
//                      maximum, like np.argmax) and resolves the XD_ROWS x XD_ROWS triangle with
//                      the same 32-row block step the window kernel uses (dp_core.cuh).
// Launches are stream-ordered, so there is no inter-CTA spinning.  Back-trace is pointer doubling.
#include "dp_core.cuh"

namespace {

constexpr int XD_ROWS = 128;           // rows per block step (4 sub-blocks of 32)
constexpr int XD_COLS = 256;           // columns per rectangle CTA
constexpr int XD_THREADS = 256;
constexpr int XD_WARPS = XD_THREADS / 32;

__global__ void gather_candidates_kernel(const int32_t *__restrict__ cand, i64 m, const i64 *__restrict__ cg,
                                         int32_t *__restrict__ L, int32_t *__restrict__ C)
{
    const i64 first = cand ? (i64)__ldg(cand) : 0;
    const i64 cg_first = __ldg(cg + first);
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (i64)gridDim.x * blockDim.x) {
        const i64 pos = cand ? (i64)__ldg(cand + k) : k;
        L[k] = (int32_t)(pos - first);
        C[k] = (int32_t)(__ldg(cg + pos) - cg_first);
    }
}

// Rectangle: rows [jb, jb+XD_ROWS) x columns chunk [c0, c1) of final columns.
// warp w: row sub-block (w & 3), column half (w >> 2) of the chunk.
template <bool AI>
__global__ void __launch_bounds__(XD_THREADS)
exact_rect_kernel(int jb, int N, int ncols, const int32_t *__restrict__ L, const int32_t *__restrict__ C,
                  const double *P, const double *__restrict__ gtab, const double *__restrict__ ltab,
                  int alpha_int, double alpha, double *part_val, int *part_arg)
{
    __shared__ int2 sLC[XD_COLS];
    __shared__ double sP[XD_COLS];
    __shared__ double sV[XD_ROWS];
    __shared__ int sA[XD_ROWS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * XD_COLS;
    const int c1 = min(c0 + XD_COLS, ncols);
    for (int i = c0 + tid; i < c1; i += XD_THREADS) {
        sLC[i - c0] = make_int2(__ldg(L + i), __ldg(C + i));
        sP[i - c0] = P[i];
    }
    __syncthreads();
    const int rs = warp & 3, half = warp >> 2;
    const int row = jb + rs * 32 + lane;
    const int j = min(row, N - 1);
    const RowConst<AI> r = make_row<AI>(__ldg(C + j), __ldg(L + j), alpha_int, alpha);
    const int nloc = c1 - c0;
    const int mid = (nloc + 1) / 2;
    const int i0 = half ? mid : 0, i1 = half ? nloc : mid;
    double best = -INFINITY;
    int arg = i0;
    sweep_columns<AI, 8>(i0, i1, sLC, sP, r, gtab, ltab, best, arg);
    if (half == 0) { sV[rs * 32 + lane] = best; sA[rs * 32 + lane] = arg; }
    __syncthreads();
    if (half == 1) {
        const double v0 = sV[rs * 32 + lane];
        const int a0 = sA[rs * 32 + lane];
        if (!(best > v0)) { best = v0; arg = a0; }     // earlier columns win ties
        part_val[(size_t)blockIdx.x * XD_ROWS + rs * 32 + lane] = best;
        part_arg[(size_t)blockIdx.x * XD_ROWS + rs * 32 + lane] = arg + c0;
    }
}

// Diagonal: merge partials, resolve rows [jb, jb+XD_ROWS) against columns [jb, row).
template <bool AI>
__global__ void __launch_bounds__(XD_THREADS)
exact_diag_kernel(int jb, int N, int nparts, const int32_t *__restrict__ L, const int32_t *__restrict__ C,
                  double *P, int *prev, const double *__restrict__ gtab, const double *__restrict__ ltab,
                  int alpha_int, double alpha, double pen, const double *part_val, const int *part_arg)
{
    __shared__ int2 sLC[XD_ROWS];
    __shared__ double sP[XD_ROWS];
    __shared__ int sPrev[XD_ROWS];
    __shared__ double sInitV[2][XD_ROWS];
    __shared__ int sInitA[2][XD_ROWS];
    __shared__ double sPartV[XD_WARPS * 32];
    __shared__ int sPartA[XD_WARPS * 32];
    __shared__ double sTri[DP_JB * DP_JB];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nrows = min(XD_ROWS, N - jb);

    if (tid < nrows) sLC[tid] = make_int2(__ldg(L + jb + tid), __ldg(C + jb + tid));
    {   // two threads per row merge the column-ordered partials
        const int row = tid & (XD_ROWS - 1), h = tid >> 7;
        const int mid = (nparts + 1) / 2;
        const int p0 = h ? mid : 0, p1 = h ? nparts : mid;
        double best = -INFINITY;
        int arg = 0;
        for (int q = p0; q < p1; ++q) {
            const double v = part_val[(size_t)q * XD_ROWS + row];
            if (v > best) { best = v; arg = part_arg[(size_t)q * XD_ROWS + row]; }
        }
        sInitV[h][row] = best;
        sInitA[h][row] = arg;
    }
    __syncthreads();

    for (int sb = 0; sb < nrows; sb += DP_JB) {
        double ib = -INFINITY;
        int ia = 0;
        if (tid < 32 && sb + lane < nrows) {
            ib = sInitV[0][sb + lane];
            ia = sInitA[0][sb + lane];
            const double v1 = sInitV[1][sb + lane];
            if (v1 > ib) { ib = v1; ia = sInitA[1][sb + lane]; }
        }
        dp_block_step<AI, XD_WARPS, 4>(sb, nrows, 0, sLC, sP, nullptr, sPrev, sPartV, sPartA, sTri, gtab, ltab,
                                    alpha_int, alpha, pen, ib, ia, jb);
    }
    if (tid < nrows) {
        P[jb + tid] = sP[tid];
        prev[jb + tid] = sPrev[tid];
    }
}

// ---- back-trace by pointer doubling over global arrays --------------------------------------
__global__ void bt_init_kernel(const int *__restrict__ prev, i64 N, int *jump, uint8_t *mark)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (i64)gridDim.x * blockDim.x) {
        jump[k] = prev[k];
        mark[k] = (k == N - 1);
    }
}
__global__ void bt_step_kernel(const int *__restrict__ jump_in, int *__restrict__ jump_out, uint8_t *mark, i64 N)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (i64)gridDim.x * blockDim.x) {
        const int jk = jump_in[k];
        if (mark[k]) mark[jk] = 1;
        jump_out[k] = jump_in[jk];
    }
}
__global__ void bt_scatter_kernel(const uint8_t *__restrict__ mark, i64 N, const int32_t *__restrict__ cand,
                                  uint32_t *keepbits)
{
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (i64)gridDim.x * blockDim.x) {
        if (mark[k]) {
            const i64 pos = cand ? (i64)__ldg(cand + k) : k;
            atomicOr(keepbits + (pos >> 5), 1u << (pos & 31));
        }
    }
}

template <bool AI>
__global__ void suffix_row_kernel(i64 stop, const int32_t *__restrict__ cand, const i64 *__restrict__ cg,
                                  const double *__restrict__ gtab, const double *__restrict__ ltab,
                                  int alpha_int, double alpha, double *__restrict__ out)
{
    const i64 first = cand ? (i64)__ldg(cand) : 0;
    const i64 cg_first = __ldg(cg + first);
    const i64 pj = cand ? (i64)__ldg(cand + stop) : stop;
    const RowConst<AI> r = make_row<AI>((int)(__ldg(cg + pj) - cg_first), (int)(pj - first), alpha_int, alpha);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < stop; i += (i64)gridDim.x * blockDim.x) {
        const i64 pi = cand ? (i64)__ldg(cand + i) : i;
        out[i] = self_score<AI>((int)(__ldg(cg + pi) - cg_first), (int)(pi - first), r, gtab, ltab);
    }
}

inline unsigned grid_for(pasio_ctx *ctx, i64 n, int threads)
{
    i64 g = (n + threads - 1) / threads;
    i64 cap = (i64)ctx->sm_count * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace

int launch_gather_candidates(pasio_ctx *ctx)
{
    const i64 m = ctx->m;
    PASIO_TRY(pasio_reserve(ctx, ctx->dpL, (size_t)m * 4));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpC, (size_t)m * 4));
    gather_candidates_kernel<<<grid_for(ctx, m, 256), 256, 0, ctx->stream>>>(
        cur_cand(ctx), m, ctx->cg.as<i64>(), ctx->dpL.as<int32_t>(), ctx->dpC.as<int32_t>());
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

template <bool AI>
static int run_exact_dp(pasio_ctx *ctx, i64 N)
{
    const int32_t *L = ctx->dpL.as<int32_t>();
    const int32_t *C = ctx->dpC.as<int32_t>();
    double *P = ctx->dpP.as<double>();
    int *prev = ctx->dpPrev.as<int>();
    const double *gtab = ctx->tab[AI ? PASIO_TAB_LGAMMA : PASIO_TAB_LGAMMA_ALPHA].as<double>();
    const double *ltab = ctx->tab[PASIO_TAB_LOG].as<double>();
    double *pv = ctx->dpPart.as<double>();
    int *pa = ctx->dpPartArg.as<int>();
    const int alpha_int = (int)ctx->alpha_int;
    // row 0: prefix_scores[0] = 0, previous_splits[0] = 0 (square_splitter.py:72,78)
    CUDA_TRY(ctx, cudaMemsetAsync(P, 0, 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(prev, 0, 4, ctx->stream));
    i64 launches = 0;
    TimingScope ts(ctx, TF_EXACT_DP, 0);
    for (i64 jb = 1; jb < N; jb += XD_ROWS) {
        const int nparts = (int)((jb + XD_COLS - 1) / XD_COLS);
        exact_rect_kernel<AI><<<nparts, XD_THREADS, 0, ctx->stream>>>((int)jb, (int)N, (int)jb, L, C, P, gtab, ltab,
                                                                     alpha_int, ctx->alpha, pv, pa);
        exact_diag_kernel<AI><<<1, XD_THREADS, 0, ctx->stream>>>((int)jb, (int)N, nparts, L, C, P, prev, gtab, ltab,
                                                                alpha_int, ctx->alpha, ctx->pen, pv, pa);
        launches += 2;
    }
    ctx->fam_launches[TF_EXACT_DP] += launches;
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_exact_dp(pasio_ctx *ctx, i64 N)
{
    PASIO_TRY(pasio_reserve(ctx, ctx->dpP, (size_t)N * 8));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpPrev, (size_t)N * 4));
    const i64 max_parts = (N + XD_COLS - 1) / XD_COLS + 1;
    PASIO_TRY(pasio_reserve(ctx, ctx->dpPart, (size_t)max_parts * XD_ROWS * 8));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpPartArg, (size_t)max_parts * XD_ROWS * 4));
    return ctx->alpha_is_int ? run_exact_dp<true>(ctx, N) : run_exact_dp<false>(ctx, N);
}

int launch_backtrace_mark(pasio_ctx *ctx, i64 N)
{
    PASIO_TRY(pasio_reserve(ctx, ctx->dpJump, (size_t)N * 8));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpMark, (size_t)N));
    int *ja = ctx->dpJump.as<int>();
    int *jb = ja + N;
    uint8_t *mark = ctx->dpMark.as<uint8_t>();
    const unsigned g = grid_for(ctx, N, 256);
    bt_init_kernel<<<g, 256, 0, ctx->stream>>>(ctx->dpPrev.as<int>(), N, ja, mark);
    for (i64 reach = 1; reach < N; reach <<= 1) {
        bt_step_kernel<<<g, 256, 0, ctx->stream>>>(ja, jb, mark, N);
        int *t = ja; ja = jb; jb = t;
    }
    const size_t bit_bytes = (size_t)((ctx->n + 1 + 31) / 32 + 2) * 4;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->keepbits.p, 0, bit_bytes, ctx->stream));
    bt_scatter_kernel<<<g, 256, 0, ctx->stream>>>(mark, N, cur_cand(ctx), ctx->keepbits.as<uint32_t>());
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_suffix_row(pasio_ctx *ctx, i64 stop, double *d_out)
{
    const double *gtab = ctx->tab[ctx->alpha_is_int ? PASIO_TAB_LGAMMA : PASIO_TAB_LGAMMA_ALPHA].as<double>();
    const double *ltab = ctx->tab[PASIO_TAB_LOG].as<double>();
    const unsigned g = grid_for(ctx, stop, 256);
    TimingScope ts(ctx, TF_SCORE);
    if (ctx->alpha_is_int)
        suffix_row_kernel<true><<<g, 256, 0, ctx->stream>>>(stop, cur_cand(ctx), ctx->cg.as<i64>(), gtab, ltab,
                                                           (int)ctx->alpha_int, ctx->alpha, d_out);
    else
        suffix_row_kernel<false><<<g, 256, 0, ctx->stream>>>(stop, cur_cand(ctx), ctx->cg.as<i64>(), gtab, ltab,
                                                            (int)ctx->alpha_int, ctx->alpha, d_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}
