// K4: all windows of one sliding-window round in ONE launch, one CTA per window.
//
// Replaces, per window, the chain
//   SlidingWindowReducer.reduce_candidates_in_window   splitters/sliding_window_reducer.py:10-18
//   -> NotConstantReducer / NotZeroReducer              splitters/constants_reducer.py:5-21
//   -> SquareSplitter.split_without_normalizations      splitters/square_splitter.py:67-100
//   -> collect_split_points                              splitters/square_splitter.py:102-109
//   -> set.update(...)                                   splitters/sliding_window_reducer.py:25
// (paths under /root/reference/src/pasio/).  In the flat formulation a window needs
// only two integer vectors: positions L and global prefix sums C of its candidates, re-based to
// the window's first candidate exactly like counts[start:stop] / candidates - start.
//
// Per CTA: (A) warp-ballot stream compaction of the window's candidates (constraint filter)
// into shared memory, (B) the DP in 32-row block steps (dp_core.cuh), (C) back-trace by pointer
// doubling and an atomicOr scatter of the survivors into the position bitmap.
// CTAs are persistent and pull windows from an atomic counter (window cost varies as N^2).
#include "dp_core.cuh"
#include <cstdlib>

namespace {

constexpr int WD_THREADS = 256;
constexpr int WD_WARPS = WD_THREADS / 32;

struct WinDpParams {
    WinGeom geom;
    i64 nwin;
    const int32_t *cand;        // nullptr: all positions
    const i64 *cg;
    const uint32_t *cpbits;
    uint32_t *keepbits;
    const double *gtab;
    const double *ltab;
    int constraint;
    int alpha_int;
    double alpha;
    double pen;
    int cap;                    // max candidates in a window
    u64 *cells;
    unsigned *work_counter;
};

__host__ __device__ inline size_t window_smem_bytes(int cap)
{
    const size_t capr = (size_t)((cap + 31) & ~31);
    return capr * 16                // sCol (L, C, P)
           + WD_WARPS * 32 * 8      // sPartV
           + DP_JB * DP_JB * 8      // sTri
           + WD_WARPS * 32 * 4      // sPartA
           + 16 * 4                 // sMisc
           + capr * 2 * 2           // sPrev, sJump (back-trace ping-pong)
           + capr;                  // sMark
}

template <bool AI, int U, int RPL>
__global__ void __launch_bounds__(WD_THREADS, 3)
window_dp_kernel(WinDpParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int capr = (p.cap + 31) & ~31;
    ColRec *sCol = reinterpret_cast<ColRec *>(smem);
    double *sPartV = reinterpret_cast<double *>(sCol + capr);
    double *sTri = sPartV + WD_WARPS * 32;
    int *sPartA = reinterpret_cast<int *>(sTri + DP_JB * DP_JB);
    int *sMisc = sPartA + WD_WARPS * 32;
    unsigned short *sPrev = reinterpret_cast<unsigned short *>(sMisc + 16);
    unsigned short *sJump = sPrev + capr;
    unsigned char *sMark = reinterpret_cast<unsigned char *>(sJump + capr);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    while (true) {
        if (tid == 0) sMisc[0] = (int)atomicAdd(p.work_counter, 1u);
        __syncthreads();
        const i64 w = (unsigned)sMisc[0];
        if (w >= p.nwin) break;

        // ---- (A) candidates of the window, filtered, re-based ---------------------------------
        i64 st, en;
        window_range(p.geom, w, st, en);
        const int nq = (int)(en - st);
        const i64 first = p.cand ? (i64)__ldg(p.cand + st) : st;
        const i64 last = p.cand ? (i64)__ldg(p.cand + en - 1) : en - 1;
        const i64 cg_first = __ldg(p.cg + first);
        const bool all_zero = (p.constraint == PASIO_CONSTRAINT_ZEROS) && (__ldg(p.cg + last) == cg_first);
        int count = 0;
        for (int base = 0; base < nq; base += WD_THREADS) {
            const int q = base + tid;
            i64 pos = 0;
            bool take = false;
            if (q < nq) {
                pos = p.cand ? (i64)__ldg(p.cand + st + q) : st + q;
                if (q == 0 || q == nq - 1 || p.constraint == PASIO_CONSTRAINT_NONE) take = true;
                else if (p.constraint == PASIO_CONSTRAINT_CONSTANTS) take = bit_test(p.cpbits, pos);
                else take = !all_zero;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (lane == 0) sMisc[4 + warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < WD_WARPS; ++w2) {
                const int c = sMisc[4 + w2];
                if (w2 < warp) woff += c;
                tot += c;
            }
            if (take) {
                const int k = count + woff + __popc(bal & ((1u << lane) - 1u));
                sCol[k].L = (int)(pos - first);
                sCol[k].C = (int)(__ldg(p.cg + pos) - cg_first);
            }
            count += tot;
            __syncthreads();
        }
        const int N = count;

        // ---- (B) DP ---------------------------------------------------------------------------
        if (tid == 0) { sCol[0].P = 0.0; sPrev[0] = 0; }
        __syncthreads();
        for (int jb = 1; jb < N; jb += DP_JB)
            dp_block_step<AI, WD_WARPS, U, RPL>(jb, N, 0, sCol, sPrev, nullptr, sPartV, sPartA, sTri,
                                                p.gtab, p.ltab, p.alpha_int, p.alpha, p.pen, -INFINITY, 0, 0);

        // ---- (C) back-trace by pointer doubling, scatter survivors ----------------------------
        for (int k = tid; k < N; k += WD_THREADS) sMark[k] = (k == N - 1);
        __syncthreads();
        unsigned short *ja = sPrev, *jb2 = sJump;
        for (int reach = 1; reach < N; reach <<= 1) {
            // nodes within `reach` hops of the end are marked; ja[k] is the node 'reach' hops before k
            for (int k = tid; k < N; k += WD_THREADS)
                if (sMark[k]) sMark[ja[k]] = 1;
            for (int k = tid; k < N; k += WD_THREADS) jb2[k] = ja[ja[k]];
            __syncthreads();
            unsigned short *t = ja; ja = jb2; jb2 = t;
        }
        for (int k = tid; k < N; k += WD_THREADS) {
            if (sMark[k]) {
                const i64 pos = first + sCol[k].L;
                atomicOr(p.keepbits + (pos >> 5), 1u << (pos & 31));
            }
        }
        if (tid == 0) atomicAdd(p.cells, (u64)N * (u64)(N - 1) / 2);
        __syncthreads();
    }
}

}  // namespace

int window_dp_max_candidates(pasio_ctx *ctx)
{
    int cap = 32;
    while (window_smem_bytes(cap + 32) <= (size_t)ctx->smem_optin) cap += 32;
    return cap;
}

int launch_window_dp(pasio_ctx *ctx, i64 nwin, int wsize, int wshift, int constraint)
{
    WinDpParams p;
    p.geom = make_geom(ctx, wsize, wshift);
    p.nwin = nwin;
    p.cand = cur_cand(ctx);
    p.cg = ctx->cg.as<i64>();
    p.cpbits = ctx->cpbits.as<uint32_t>();
    p.keepbits = ctx->keepbits.as<uint32_t>();
    p.gtab = ctx->tab[ctx->alpha_is_int ? PASIO_TAB_LGAMMA : PASIO_TAB_LGAMMA_ALPHA].as<double>();
    p.ltab = ctx->tab[PASIO_TAB_LOG].as<double>();
    p.constraint = constraint;
    p.alpha_int = (int)ctx->alpha_int;
    p.alpha = ctx->alpha;
    p.pen = ctx->pen;
    i64 cap = (i64)wsize + 1;
    if (cap > ctx->m) cap = ctx->m;
    if (cap > 65535 || window_smem_bytes((int)cap) > (size_t)ctx->smem_optin)
        return pasio_fail(ctx, PASIO_E_TOO_LARGE, "window of %lld candidates does not fit one CTA's shared memory (max %d)",
                          (long long)cap, window_dp_max_candidates(ctx));
    p.cap = (int)cap;
    p.cells = ctx->scalars.as<u64>() + 10;
    p.work_counter = ctx->scalars.as<unsigned>() + 2 * 11;   // scalars[11]
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->scalars.as<u64>() + 10, 0, 16, ctx->stream));

    const size_t smem = window_smem_bytes(p.cap);
    // rows per lane in the rectangle sweep (see dp_block_step); PASIO_WD_RPL overrides for experiments
    static const int rpl = getenv("PASIO_WD_RPL") ? atoi(getenv("PASIO_WD_RPL")) : 2;
    void (*kern)(WinDpParams);
    if (ctx->alpha_is_int) kern = rpl == 1 ? window_dp_kernel<true, 8, 1> : window_dp_kernel<true, 4, 2>;
    else kern = rpl == 1 ? window_dp_kernel<false, 8, 1> : window_dp_kernel<false, 4, 2>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WD_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    i64 grid = (i64)ctx->sm_count * per_sm;     // persistent CTAs: one resident wave
    if (grid > nwin) grid = nwin;
    if (grid < 1) grid = 1;
    {
        TimingScope ts(ctx, TF_WINDOW_DP);
        kern<<<(unsigned)grid, WD_THREADS, smem, ctx->stream>>>(p);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}
