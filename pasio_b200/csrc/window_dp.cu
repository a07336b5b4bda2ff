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
constexpr int PR_NEAR = 64;     // pruned path: columns this close to the row block are always evaluated
constexpr int PR_FB = 8;        // pruned path: far columns are bounded in blocks of 8

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
    u64 *cells;                 // algorithmic cells N(N-1)/2
    u64 *cells_skipped;         // cells proven irrelevant by the far-column bound (0 without pruning)
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
           + capr                   // sMark
           + (capr / PR_FB) * 8     // sBMax: max P of every 8-column block (pruned path)
           + 32 * 8 * 2 + 32 * 4    // sLB, sFarV, sFarA
           + 4 * 8;                 // sScal
}

// ---- exact pruning of far columns (branch and bound) -----------------------------------------
// For row j and a block I of consecutive columns [i0, i1):  every cell value
//     t_ij = (G[s_ij] - s_ij * Lg[len_ij]) + P_i ,  s_ij = S_j - C_i ,  len_ij = L_j - L_i
// obeys   t_ij <= max(F(s_lo, len_lo), F(s_hi, len_lo)) + max_{i in I} P_i + delta   where
// F(s, len) = G[s] - s*Lg[len], s_lo/s_hi are the block's extreme counts and len_lo its shortest
// length: F decreases in len (s >= 0, log non-decreasing) and lgamma(s) - s*c is convex in s, so over
// the block it is largest at an end point; delta covers the table and rounding errors (2^-44 of the
// window's largest magnitudes, >1000x the worst case, <1e-6 in absolute terms).
// A block whose bound is strictly below a LOWER bound of the row's maximum cannot hold the
// arg-max nor tie with it, so skipping it leaves P, prev and the back-trace bit-identical.
// The lower bound is the row's maximum over the PR_NEAR nearest columns plus the triangle, obtained
// by running the block chain once on those columns only (a feasible segmentation, hence <= optimum).
// Surviving blocks are evaluated exactly, cell by cell, in the reference's operation order.
template <bool AI>
__device__ __forceinline__ void lex_max(double &best, int &arg, double v, int a)
{
    if (v > best || (v == best && a < arg)) { best = v; arg = a; }
}

template <bool AI>
__device__ __forceinline__ u64 far_pass(int jb, int N, int nfar, const ColRec *sCol, const double *sBMax,
                                        const double *sLB, double *sFarV, int *sFarA, double delta,
                                        const double *__restrict__ gtab, const double *__restrict__ ltab,
                                        int alpha_int, double alpha)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rr = lane & 3, cc = lane >> 2;              // bound phase: 4 rows x 8 column blocks
    const int sub = lane >> 3, l8 = lane & 7;             // exact phase: 4 surviving blocks x 8 columns
    const int r = 4 * warp + rr;
    const bool row_ok = jb + r < N;
    const ColRec me = sCol[min(jb + r, N - 1)];
    const RowConst<AI> rc = make_row<AI>(me.C, me.L, alpha_int, alpha);
    const double lb = row_ok ? sLB[r] : INFINITY;
    u64 skipped = 0;

    // column 0 is not part of any 8-block: evaluate it exactly and seed the far result with it
    if (cc == 0 && row_ok) {
        const ColRec a = sCol[0];
        sFarV[r] = __dadd_rn(self_score<AI>(a.C, a.L, rc, gtab, ltab), a.P);
        sFarA[r] = 0;
    }
    __syncwarp();

    for (int cb0 = 0; cb0 < nfar; cb0 += 8) {
        const int b = cb0 + cc;
        bool surv = false;
        if (row_ok && b < nfar) {
            const ColRec a = sCol[1 + PR_FB * b];                  // first column: largest count, longest length
            const ColRec z = sCol[PR_FB * b + PR_FB];              // last column: smallest count, shortest length
            const int x_hi = rc.cjx - a.C, x_lo = rc.cjx - z.C;
            const double lg = __ldg(ltab + (rc.lj - z.L));
            const double s_hi = AI ? u32_to_double(x_hi) : __dsub_rn(rc.aj, u32_to_double(a.C));
            const double s_lo = AI ? u32_to_double(x_lo) : __dsub_rn(rc.aj, u32_to_double(z.C));
            const double f_hi = __dsub_rn(__ldg(gtab + x_hi), __dmul_rn(s_hi, lg));
            const double f_lo = __dsub_rn(__ldg(gtab + x_lo), __dmul_rn(s_lo, lg));
            const double ub = fmax(f_hi, f_lo) + sBMax[b] + delta;
            surv = !(ub < lb);                                      // NaN keeps the block
            if (!surv) skipped += PR_FB;
        }
        unsigned mask = __ballot_sync(0xffffffffu, surv);
        while (mask) {
            // the four lowest survivors are evaluated together, 8 lanes (columns) each
            int sel = -1;
            unsigned m = mask;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int bit = m ? (__ffs(m) - 1) : -1;
                if (k == sub) sel = bit;
                m &= m - 1;                                         // 0 & anything stays 0
            }
            mask = m;
            double t = -INFINITY;
            int col = 0x7fffffff, row2 = 0;
            if (sel >= 0) {
                row2 = 4 * warp + (sel & 3);
                const int b2 = cb0 + (sel >> 2);
                col = 1 + PR_FB * b2 + l8;
                const ColRec rowrec = sCol[jb + row2];
                const RowConst<AI> r2 = make_row<AI>(rowrec.C, rowrec.L, alpha_int, alpha);
                const ColRec a = sCol[col];
                t = __dadd_rn(self_score<AI>(a.C, a.L, r2, gtab, ltab), a.P);
            }
#pragma unroll
            for (int off = 4; off > 0; off >>= 1) {
                const double ot = __shfl_xor_sync(0xffffffffu, t, off);
                const int oc = __shfl_xor_sync(0xffffffffu, col, off);
                lex_max<AI>(t, col, ot, oc);
            }
            // two survivors may belong to the same row: apply the four results one after another
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (sub == k && l8 == 0 && sel >= 0) {
                    double bv = sFarV[row2];
                    int ba = sFarA[row2];
                    lex_max<AI>(bv, ba, t, col);
                    sFarV[row2] = bv;
                    sFarA[row2] = ba;
                }
                __syncwarp();
            }
        }
    }
    return skipped;
}

template <bool AI, int U, int RPL, bool PRUNE>
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
    double *sBMax = reinterpret_cast<double *>(sMark + capr);      // capr is a multiple of 32: stays 8-byte aligned
    double *sLB = sBMax + capr / PR_FB;
    double *sFarV = sLB + 32;
    double *sScal = sFarV + 32;                                     // [0] magnitude of the window's largest self score, [1] max |P|
    int *sFarA = reinterpret_cast<int *>(sScal + 4);

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
        u64 skipped = 0;
        if (PRUNE && tid == 0) {
            // largest |G| + s*Lg any cell of this window can reach (both monotone): the scale of delta
            const ColRec z = sCol[N - 1];
            const int x = z.C + (AI ? p.alpha_int : 0);
            sScal[0] = fabs(__ldg(p.gtab + x)) + ((double)z.C + p.alpha) * fabs(__ldg(p.ltab + z.L)) + 1.0;
            sScal[1] = 0.0;
        }
        for (int jb = 1; jb < N; jb += DP_JB) {
            constexpr int NQ = WD_WARPS / (DP_JB / (DP_RPW * RPL));
            const int near_lo = jb - PR_NEAR;
            if (!PRUNE || near_lo < 1 + PR_FB) {
                dp_block_step<AI, WD_WARPS, U, RPL>(jb, N, 0, sCol, sPrev, nullptr, sPartV, sPartA, sTri,
                                                    p.gtab, p.ltab, p.alpha_int, p.alpha, p.pen, -INFINITY, 0, 0);
            } else {
                // (1) nearest columns + triangle, (2) provisional chain -> lower bounds
                block_rect_tri<AI, WD_WARPS, U, RPL>(jb, N, near_lo, sCol, sPartV, sPartA, sTri, p.gtab, p.ltab,
                                                     p.alpha_int, p.alpha);
                if (warp == 0)
                    block_chain<NQ>(jb, N, sCol, sPrev, nullptr, sPartV, sPartA, sTri, p.pen, -INFINITY, 0, 0, sLB);
                __syncthreads();
                // (3) far columns [1, near_lo): bound blocks of 8, evaluate the survivors exactly
                const double delta = ldexp(sScal[0] + sScal[1], -44);
                skipped += far_pass<AI>(jb, N, (near_lo - 1) / PR_FB, sCol, sBMax, sLB, sFarV, sFarA, delta,
                                          p.gtab, p.ltab, p.alpha_int, p.alpha);
                __syncthreads();
                // (4) final chain: far result first (smaller columns), then the near partials, then the triangle
                if (warp == 0)
                    block_chain<NQ>(jb, N, sCol, sPrev, nullptr, sPartV, sPartA, sTri, p.pen,
                                    jb + lane < N ? sFarV[lane] : -INFINITY, jb + lane < N ? sFarA[lane] : 0, 0, nullptr);
                __syncthreads();
            }
            if (PRUNE && warp == 0) {
                // per-8-column maxima of the finished rows and the running max |P| (scale of delta)
                const double pv = jb + lane < N ? sCol[jb + lane].P : -INFINITY;
                double mx = pv, ab = jb + lane < N ? fabs(pv) : 0.0;
#pragma unroll
                for (int off = 4; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ab = fmax(ab, __shfl_xor_sync(0xffffffffu, ab, off));
                if ((lane & 7) == 0) sBMax[(jb - 1) / PR_FB + (lane >> 3)] = mx;
                if (lane == 0) sScal[1] = fmax(sScal[1], ab);
            }
        }

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
        if (PRUNE) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) skipped += __shfl_xor_sync(0xffffffffu, skipped, off);
            if (lane == 0 && skipped) atomicAdd(p.cells_skipped, skipped);
        }
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
    p.cells_skipped = ctx->scalars.as<u64>() + 12;
    p.work_counter = ctx->scalars.as<unsigned>() + 2 * 11;   // scalars[11]
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->scalars.as<u64>() + 10, 0, 24, ctx->stream));

    const size_t smem = window_smem_bytes(p.cap);
    // PASIO_WD_PRUNE=0 disables the exact far-column pruning (experiments / cross-checks)
    static const int prune_env = getenv("PASIO_WD_PRUNE") ? atoi(getenv("PASIO_WD_PRUNE")) : 1;
    const bool prune = prune_env != 0 && ctx->alpha >= 0.0009765625;   // tiny alpha: lgamma(alpha) dwarfs the delta scale; alpha = 0: G[0] = inf
    void (*kern)(WinDpParams);
    if (ctx->alpha_is_int) kern = prune ? window_dp_kernel<true, 4, 2, true> : window_dp_kernel<true, 4, 2, false>;
    else kern = prune ? window_dp_kernel<false, 4, 2, true> : window_dp_kernel<false, 4, 2, false>;
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
