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
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int WD_THREADS = 256;

// Development-only phase timers (make PROF=1): cycles seen by thread 0 of every CTA, summed per phase.
#ifdef PASIO_WD_PROF
__device__ unsigned long long g_wd_prof[24];
struct ProfAcc { long long t0; long long a0, a1, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12; };
#define PROF_DECL ProfAcc prof = {clock64(), 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_T(i) do { asm volatile("" ::: "memory"); const long long t1__ = clock64(); prof.a##i += t1__ - prof.t0; prof.t0 = t1__; asm volatile("" ::: "memory"); } while (0)
#define PROF_FLUSH do { if (threadIdx.x == 0) { \
    atomicAdd(g_wd_prof + 0, (unsigned long long)prof.a0); atomicAdd(g_wd_prof + 1, (unsigned long long)prof.a1); \
    atomicAdd(g_wd_prof + 3, (unsigned long long)prof.a3); atomicAdd(g_wd_prof + 4, (unsigned long long)prof.a4); \
    atomicAdd(g_wd_prof + 5, (unsigned long long)prof.a5); atomicAdd(g_wd_prof + 6, (unsigned long long)prof.a6); \
    atomicAdd(g_wd_prof + 7, (unsigned long long)prof.a7); atomicAdd(g_wd_prof + 8, (unsigned long long)prof.a8); \
    atomicAdd(g_wd_prof + 9, (unsigned long long)prof.a9); atomicAdd(g_wd_prof + 10, (unsigned long long)prof.a10); \
    atomicAdd(g_wd_prof + 11, (unsigned long long)prof.a11); atomicAdd(g_wd_prof + 12, (unsigned long long)prof.a12); } \
    prof.a0 = prof.a1 = prof.a3 = prof.a4 = prof.a5 = prof.a6 = prof.a7 = prof.a8 = prof.a9 = prof.a10 = prof.a11 = prof.a12 = 0; } while (0)
#define PROF_ARGS , ProfAcc &prof
#define PROF_PASS , prof
#else
#define PROF_DECL
#define PROF_T(i) do { } while (0)
#define PROF_FLUSH do { } while (0)
#define PROF_ARGS
#define PROF_PASS
#endif
constexpr int WD_WARPS = WD_THREADS / 32;
constexpr int PR_CB = 32;       // pruned path: far columns are first bounded in blocks of 32 ...
constexpr int PR_FB = 8;        // ... and the surviving blocks again in sub-blocks of 8

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
#ifdef PASIO_WD_EXP
    int exp_skip;               // timing experiments only (results are wrong): 1 far pass, 2 record fit, 4 chain, 8 near rectangle
#endif
};
#ifdef PASIO_WD_EXP
#define EXP_SKIP(bit) (p.exp_skip & (bit))
#else
#define EXP_SKIP(bit) false
#endif

// One finished block of 32 columns [1+32b, 32+32b], as the far pass sees it.
struct __align__(16) CoarseRec {
    int c_first, c_last, l_first, l_last;   // C and L of its first / last column
    double a, b;                            // tilt: P_i + a*C_i + b*L_i is nearly constant over the block
    double mpt;                             // max_i (P_i + a*C_i + b*L_i) over the block, raised by the tilt's rounding slack
    double mpt8[4];                         // the same over each 8-column sub-block
    double pad;
};
static_assert(sizeof(CoarseRec) == 80, "CoarseRec layout");

// One row of the current 32-row block step: a lower bound of its maximum, its C and L.
struct __align__(16) RowLB {
    double lb;
    int C;
    int L;
};

__host__ __device__ inline size_t window_smem_bytes(int cap)
{
    const size_t capr = (size_t)((cap + 31) & ~31);
    return capr * 16                // sCol (L, C, P)
           + WD_WARPS * 32 * 8      // sPartV
           + DP_JB * DP_JB * 8      // sTri            (back-trace: sMark lives here, needs capr <= 8192)
           + WD_WARPS * 32 * 4      // sPartA
           + 16 * 4                 // sMisc
           + capr * 2               // sPrev
           + (capr / PR_CB) * sizeof(CoarseRec)   // sCoarse  (back-trace: sJump lives here, capr*2 bytes)
           + WD_WARPS * 32 * sizeof(RowLB)        // sRow: every warp keeps its own copy of the 32 row records
           + WD_WARPS * 32 * 8 + WD_WARPS * 32 * 4   // sFarV, sFarA: per-warp far results of the 32 rows
           + 4 * 8;                 // sScal
}

// ---- exact pruning of far columns (branch and bound) -----------------------------------------
// For rows j in [j0, j1] and columns i in [i0, i1] every cell value
//     t_ij = F(u_ij, len_ij) + P_i ,  F(u, len) = G[u (+alpha)] - s*Lg[len],  u_ij = C_j - C_i,  len_ij = L_j - L_i
// satisfies, for ANY real a, b (the "tilt"),
//     t_ij - LB_j = [P_i + a*C_i + b*L_i] + [F(u_ij, len_ij) + a*u_ij + b*len_ij] - [LB_j + a*C_j + b*L_j]
//                <= max_i [P_i + a*C_i + b*L_i]  +  max_box [F(u, len) + a*u + b*len]  -  min_j [LB_j + a*C_j + b*L_j]
// where the box is [C_j0 - C_i1, C_j1 - C_i0] x [L_j0 - L_i1, L_j1 - L_i0].  F + a*u + b*len is convex in u for fixed
// len (lgamma is convex, the rest is linear) and convex in len for fixed u (-s*log(len + beta) with s >= 0), so
// its maximum over the box is at one of the 4 corners.  With a, b fitted to the block's own P (least squares: along
// an optimal path P is close to linear in (C, L)) the first bracket hardly varies and the bound is tight to second
// order; without the tilt it is loose by the block's whole range of P.  If the right-hand side (plus delta, which
// covers table and rounding errors: 2^-44 of the window's largest magnitudes, > 50x the worst case, < 1e-6 in
// absolute terms) is negative, no cell of the rectangle reaches the lower bound LB_j of its row's maximum: it can
// hold neither the arg-max nor a tie, so skipping it leaves P, prev and the back-trace bit-identical.
// LB_j = max( best over the nearest 32 columns, the "split at every candidate" path through the block's own rows ):
// both are feasible segmentations; the path value is formed with a parallel prefix sum and lowered by delta to
// stay below the sequentially rounded value the chain would produce.
// Two levels: 32 rows x 32 columns first (one lane per rectangle), the survivors again as 4 rows x 8 columns
// (one warp per surviving rectangle); what survives both is evaluated exactly, cell by cell, in the reference's
// operation order.
template <bool AI>
__device__ __forceinline__ double tilted_box_max(int u_lo, int u_hi, int len_lo, int len_hi, double a, double b,
                                                 const double *__restrict__ gtab, const double *__restrict__ ltab,
                                                 int alpha_int, double alpha)
{
    const double g_lo = __ldg(gtab + (AI ? u_lo + alpha_int : u_lo)), g_hi = __ldg(gtab + (AI ? u_hi + alpha_int : u_hi));
    const double l_lo = __ldg(ltab + len_lo), l_hi = __ldg(ltab + len_hi);
    const double ud_lo = u32_to_double(u_lo), ud_hi = u32_to_double(u_hi);
    const double s_lo = ud_lo + alpha, s_hi = ud_hi + alpha;
    const double ta_lo = a * ud_lo, ta_hi = a * ud_hi;
    const double tb_lo = b * u32_to_double(len_lo), tb_hi = b * u32_to_double(len_hi);
    const double f00 = (g_lo - s_lo * l_lo) + (ta_lo + tb_lo);
    const double f01 = (g_lo - s_lo * l_hi) + (ta_lo + tb_hi);
    const double f10 = (g_hi - s_hi * l_lo) + (ta_hi + tb_lo);
    const double f11 = (g_hi - s_hi * l_hi) + (ta_hi + tb_hi);
    return fmax(fmax(f00, f01), fmax(f10, f11));
}

// Far columns [1, 1 + 32*nfar) of one 32-row block step.  Coarse block cb is bounded by warp cb % 8, lane cb / 8.
// Every warp keeps a running (max, first arg-max) for all 32 rows (lane = row) over the cells it evaluated exactly
// and publishes it in its slice of sFarV / sFarA; warp 0 also covers column 0, which is in no block.
// Returns the number of cells skipped (per lane; the caller sums).
template <bool AI, int NQ>
__device__ __forceinline__ u64 far_pass(int jb, int N, int nfar, const ColRec *sCol, const CoarseRec *sCoarse,
                                        const double *sPartV, RowLB *sRowW, double *sFarVW, int *sFarAW,
                                        double delta, double pen,
                                        const double *__restrict__ gtab, const double *__restrict__ ltab,
                                        int alpha_int, double alpha PROF_ARGS)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrows = min(DP_JB, N - jb);

    // ---- lower bounds for all 32 rows (every warp computes them; no block-wide sync needed) ----
    const int jrow = min(jb + lane, N - 1);
    const ColRec me = sCol[jrow];
    const RowConst<AI> rme = make_row<AI>(me.C, me.L, alpha_int, alpha);
    {
        const ColRec before = sCol[jrow - 1];
        double run = self_score<AI>(before.C, before.L, rme, gtab, ltab) + (lane ? pen : 0.0);   // w(j-1, j) [+ pen of the previous row]
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, run, d);
            if (lane >= d) run += o;
        }
        double path = sCol[jb - 1].P + run - delta;      // P_{jb-1} + sum_{m<=l} w(m-1,m) + l*pen, kept below its rounded value
        if (!(path == path)) path = -INFINITY;           // inf - inf etc.: no information
        double nearbest = sPartV[lane];
#pragma unroll
        for (int q = 1; q < NQ; ++q) nearbest = fmax(nearbest, sPartV[q * 32 + lane]);
        RowLB r;
        r.lb = fmax(path, nearbest);
        r.C = me.C;
        r.L = me.L;
        sRowW[lane] = r;
    }
    __syncwarp();

    // running far result of row `lane`
    double best = -INFINITY;
    int arg = 0x7fffffff;
    if (warp == 0 && lane < nrows) {
        const ColRec a0 = sCol[0];
        best = __dadd_rn(self_score<AI>(a0.C, a0.L, rme, gtab, ltab), a0.P);
        arg = 0;
    }
    u64 skipped = 0;
    PROF_T(10);

    // ---- level 1: 32 rows x 32 columns, one lane per rectangle ----
    const RowLB rowF = sRowW[0], rowL = sRowW[nrows - 1];
    const int cb = lane * WD_WARPS + warp;
    bool surv1 = false;
    if (cb < nfar) {
        const CoarseRec *rec = sCoarse + cb;
        const int4 ends = *reinterpret_cast<const int4 *>(rec);
        const double a = rec->a, b = rec->b;
        const double m2 = tilted_box_max<AI>(rowF.C - ends.y, rowL.C - ends.x, rowF.L - ends.w, rowL.L - ends.z, a, b,
                                             gtab, ltab, alpha_int, alpha);
        double m3 = INFINITY;
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) {
            const RowLB rr = sRowW[r];
            m3 = fmin(m3, rr.lb + (a * u32_to_double(rr.C) + b * u32_to_double(rr.L)));
        }
        surv1 = !(rec->mpt + m2 - m3 + delta < 0.0);                 // NaN keeps the block
        if (!surv1) skipped += (u64)(PR_CB * nrows);
    }
    unsigned mask1 = __ballot_sync(0xffffffffu, surv1);
    PROF_T(11);

    // ---- level 2: a surviving block as 8 row groups x 4 sub-blocks of 8 columns, one lane per rectangle ----
    const int rg = lane >> 2, q = lane & 3;
    const int r0 = 4 * rg, r1 = min(r0 + 3, nrows - 1);
    const bool act2 = r0 < nrows;
    const RowLB gF = sRowW[min(r0, nrows - 1)], gL = sRowW[r1];
    while (mask1) {
        const int cb1 = (__ffs(mask1) - 1) * WD_WARPS + warp;
        mask1 &= mask1 - 1;
        const CoarseRec *rec = sCoarse + cb1;
        const double a = rec->a, b = rec->b;
        bool surv2 = false;
        if (act2) {
            const int i0 = 1 + PR_CB * cb1 + PR_FB * q;
            const ColRec cF = sCol[i0], cL = sCol[i0 + PR_FB - 1];
            const double m2 = tilted_box_max<AI>(gF.C - cL.C, gL.C - cF.C, gF.L - cL.L, gL.L - cF.L, a, b,
                                                 gtab, ltab, alpha_int, alpha);
            double m3 = INFINITY;
            for (int r = r0; r <= r1; ++r) {
                const RowLB rr = sRowW[r];
                m3 = fmin(m3, rr.lb + (a * u32_to_double(rr.C) + b * u32_to_double(rr.L)));
            }
            surv2 = !(rec->mpt8[q] + m2 - m3 + delta < 0.0);
            if (!surv2) skipped += (u64)(PR_FB * (r1 - r0 + 1));
        }
        unsigned mask2 = __ballot_sync(0xffffffffu, surv2);

        // ---- level 3: the survivors exactly; lane = (row r of the group, column c of the sub-block) ----
        const int er = lane >> 3, ec = lane & 7;
        while (mask2) {
            const int l2 = __ffs(mask2) - 1;
            mask2 &= mask2 - 1;
            const int row = 4 * (l2 >> 2) + er;                    // row of the block step
            const int col = 1 + PR_CB * cb1 + PR_FB * (l2 & 3) + ec;
            double t = -INFINITY;
            int ta = col;
            if (row < nrows) {
                const RowLB rr = sRowW[row];
                const RowConst<AI> rc = make_row<AI>(rr.C, rr.L, alpha_int, alpha);
                const ColRec cc = sCol[col];
                t = __dadd_rn(self_score<AI>(cc.C, cc.L, rc, gtab, ltab), cc.P);
            }
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, t, off);
                const int oa = __shfl_xor_sync(0xffffffffu, ta, off);
                if (ob > t || (ob == t && oa < ta)) { t = ob; ta = oa; }
            }
            // hand the 4 row results (lanes 0, 8, 16, 24) to the lanes that own those rows
            const int rel = lane - 4 * (l2 >> 2);
            const double v = __shfl_sync(0xffffffffu, t, (rel & 3) * 8);
            const int va = __shfl_sync(0xffffffffu, ta, (rel & 3) * 8);
            if (rel >= 0 && rel < 4 && (v > best || (v == best && va < arg))) { best = v; arg = va; }
        }
    }
    sFarVW[lane] = best;
    sFarAW[lane] = arg;
    PROF_T(12);
    return skipped;
}

// After the chain finished rows [jb, jb+32) (a full block of 32 columns from now on): least-squares tilt, tilted
// maxima, end points.  One warp, lane = column.
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ void build_coarse_record(int jb, const ColRec *sCol, CoarseRec *rec, double tilt_scale_c,
                                                    double tilt_scale_l)
{
    const int lane = threadIdx.x & 31;
    const ColRec me = sCol[jb + lane];
    const int c_first = __shfl_sync(0xffffffffu, me.C, 0), c_last = __shfl_sync(0xffffffffu, me.C, 31);
    const int l_first = __shfl_sync(0xffffffffu, me.L, 0), l_last = __shfl_sync(0xffffffffu, me.L, 31);
    const double p_first = __shfl_sync(0xffffffffu, me.P, 0);
    const double x = u32_to_double(me.C - c_first), y = u32_to_double(me.L - l_first), p = me.P - p_first;
    double a = 0.0, b = 0.0;                       // fit  -P ~ a*C + b*L + const
#ifndef PASIO_NO_LS
    const double inv_n = 1.0 / 32.0;
    const double sx = warp_sum(x), sy = warp_sum(y), sp = warp_sum(p);
    const double xc = x - sx * inv_n, yc = y - sy * inv_n, pc = p - sp * inv_n;
    const double cxx = warp_sum(xc * xc), cyy = warp_sum(yc * yc), cxy = warp_sum(xc * yc);
    const double cxp = warp_sum(xc * pc), cyp = warp_sum(yc * pc);
    const double det = cxx * cyy - cxy * cxy;
    if (det > 1e-9 * cxx * cyy) {
        a = -(cxp * cyy - cyp * cxy) / det;
        b = -(cyp * cxx - cxp * cxy) / det;
    } else if (cxx > 0.0) {
        a = -cxp / cxx;
    } else if (cyy > 0.0) {
        b = -cyp / cyy;
    }
    if (!(fabs(a) < 1e300) || !(fabs(b) < 1e300)) { a = 0.0; b = 0.0; }     // any finite tilt is valid; NaN / inf is not
#endif
    double m = me.P + (a * u32_to_double(me.C) + b * u32_to_double(me.L));
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    double m32 = m;
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) m32 = fmax(m32, __shfl_xor_sync(0xffffffffu, m32, off));
    const double slack = ldexp(fabs(a) * tilt_scale_c + fabs(b) * tilt_scale_l, -44);
    if ((lane & 7) == 0) rec->mpt8[lane >> 3] = m + slack;
    if (lane == 0) {
        *reinterpret_cast<int4 *>(rec) = make_int4(c_first, c_last, l_first, l_last);
        rec->a = a;
        rec->b = b;
        rec->mpt = m32 + slack;
    }
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
    CoarseRec *sCoarse = reinterpret_cast<CoarseRec *>(sPrev + capr);   // capr*2 bytes is a multiple of 16
    RowLB *sRow = reinterpret_cast<RowLB *>(sCoarse + capr / PR_CB);
    double *sFarV = reinterpret_cast<double *>(sRow + WD_WARPS * 32);
    double *sScal = sFarV + WD_WARPS * 32;                              // [0] magnitude of the window's largest self score, [1] max |P|, [2], [3] tilt scales
    int *sFarA = reinterpret_cast<int *>(sScal + 4);
    // the back-trace runs after the DP, when these are dead
    unsigned short *sJump = reinterpret_cast<unsigned short *>(sCoarse); // capr*2 bytes <= (capr/32)*80
    unsigned char *sMark = reinterpret_cast<unsigned char *>(sTri);     // capr <= 8192 bytes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    PROF_DECL;

    while (true) {
        PROF_T(9);
        PROF_FLUSH;
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
        const bool by_words = !p.cand && p.constraint == PASIO_CONSTRAINT_CONSTANTS;
        if (by_words) {
            // all positions are candidates (round 1): the survivors of the NotConstant filter are the set
            // bits of the change-point bitmap inside [first, last], plus both ends -- one thread per word
            const i64 w0 = first >> 5, w1 = last >> 5;
            const int nwords = (int)(w1 - w0 + 1);
            for (int base = 0; base < nwords; base += WD_THREADS) {
                const int t = base + tid;
                unsigned word = 0;
                if (t < nwords) {
                    word = __ldg(p.cpbits + w0 + t);
                    if (t == 0) word = (word & (0xffffffffu << (first & 31))) | (1u << (first & 31));
                    if (t == nwords - 1) word = (word & (0xffffffffu >> (31 - (last & 31)))) | (1u << (last & 31));
                }
                int incl = __popc(word);
                const int mine = incl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += o;
                }
                if (lane == 31) sMisc[4 + warp] = incl;
                __syncthreads();
                int woff = 0, tot = 0;
#pragma unroll
                for (int w2 = 0; w2 < WD_WARPS; ++w2) {
                    const int c = sMisc[4 + w2];
                    if (w2 < warp) woff += c;
                    tot += c;
                }
                int k = count + woff + incl - mine;
                const i64 pos0 = (w0 + t) << 5;
                while (word) {
                    const i64 pos = pos0 + (__ffs(word) - 1);
                    word &= word - 1;
                    sCol[k].L = (int)(pos - first);
                    sCol[k].C = (int)(__ldg(p.cg + pos) - cg_first);
                    ++k;
                }
                count += tot;
                __syncthreads();
            }
        }
        for (int base = 0; !by_words && base < nq; base += WD_THREADS) {
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
        PROF_T(0);

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
            sScal[2] = (double)z.C + p.alpha;       // |a*C| and |b*L| of a tilt stay below |a|*this and |b|*that
            sScal[3] = (double)z.L;
        }
        for (int jb = 1; jb < N; jb += DP_JB) {
            constexpr int NQ = WD_WARPS / (DP_JB / (DP_RPW * RPL));
            const int nfar = (jb - 1) / PR_CB - 1;                  // finished 32-column blocks before the nearest one
            if (!PRUNE || nfar < 1) {
                dp_block_step<AI, WD_WARPS, U, RPL>(jb, N, 0, sCol, sPrev, nullptr, sPartV, sPartA, sTri,
                                                    p.gtab, p.ltab, p.alpha_int, p.alpha, p.pen, -INFINITY, 0, 0);
                PROF_T(8);
            } else {
                // (1) nearest 32 columns + triangle self scores
                if (!EXP_SKIP(8))
                block_rect_tri<AI, WD_WARPS, 2, RPL>(jb, N, jb - PR_CB, sCol, sPartV, sPartA, sTri, p.gtab, p.ltab,
                                                     p.alpha_int, p.alpha);
                PROF_T(1);
                // (2) far columns [1, jb - 32): two levels of bounds against a lower bound of the row maxima,
                //     the survivors exactly
                const double delta = ldexp(sScal[0] + sScal[1] + fabs(p.pen) * DP_JB, -44);
                if (!EXP_SKIP(1))
                skipped += far_pass<AI, NQ>(jb, N, nfar, sCol, sCoarse, sPartV, sRow + warp * 32, sFarV + warp * 32,
                                            sFarA + warp * 32, delta, p.pen, p.gtab, p.ltab, p.alpha_int, p.alpha PROF_PASS);
                __syncthreads();
                PROF_T(3);
                // (3) chain: far results first (smaller columns; the warps' column sets interleave, hence the
                //     index-aware merge), then the near partials, then the triangle
                if (warp == 0 && !EXP_SKIP(4)) {
                    double fbest = -INFINITY;
                    int farg = 0;
#pragma unroll
                    for (int w2 = 0; w2 < WD_WARPS; ++w2) {
                        const double v = sFarV[w2 * 32 + lane];
                        const int a = sFarA[w2 * 32 + lane];
                        if (v > fbest || (v == fbest && a < farg)) { fbest = v; farg = a; }
                    }
                    block_chain<NQ>(jb, N, sCol, sPrev, nullptr, sPartV, sPartA, sTri, p.pen,
                                    jb + lane < N ? fbest : -INFINITY, jb + lane < N ? farg : 0, 0, nullptr);
                }
                PROF_T(4);
                __syncthreads();
                PROF_T(5);
            }
            if (PRUNE && warp == 0) {
                // the finished rows become a block of 32 columns; running max |P| (scale of delta)
                if (jb + DP_JB <= N && !EXP_SKIP(2)) build_coarse_record(jb, sCol, sCoarse + (jb - 1) / PR_CB, sScal[2], sScal[3]);
                double ab = jb + lane < N ? fabs(sCol[jb + lane].P) : 0.0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ab = fmax(ab, __shfl_xor_sync(0xffffffffu, ab, off));
                if (lane == 0) sScal[1] = fmax(sScal[1], ab);
            }
            PROF_T(6);
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
        PROF_T(7);
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
    while (cap + 32 <= 8192 && window_smem_bytes(cap + 32) <= (size_t)ctx->smem_optin) cap += 32;
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
    if (cap > 8192 || window_smem_bytes((int)cap) > (size_t)ctx->smem_optin)
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
#if defined(PASIO_WD_EXP) || defined(PASIO_WD_PROF)
    if (getenv("PASIO_WD_CTAS") && atoi(getenv("PASIO_WD_CTAS")) < per_sm) per_sm = atoi(getenv("PASIO_WD_CTAS"));
#endif
#ifdef PASIO_WD_EXP
    p.exp_skip = getenv("PASIO_WD_SKIP") ? atoi(getenv("PASIO_WD_SKIP")) : 0;
    if (getenv("PASIO_WD_CTAS") && atoi(getenv("PASIO_WD_CTAS")) < per_sm) per_sm = atoi(getenv("PASIO_WD_CTAS"));
#endif
    i64 grid = (i64)ctx->sm_count * per_sm;     // persistent CTAs: one resident wave
    if (grid > nwin) grid = nwin;
    if (grid < 1) grid = 1;
    {
        TimingScope ts(ctx, TF_WINDOW_DP);
        kern<<<(unsigned)grid, WD_THREADS, smem, ctx->stream>>>(p);
    }
    CUDA_TRY(ctx, cudaGetLastError());
#ifdef PASIO_WD_PROF
    {
        cudaStreamSynchronize(ctx->stream);
        unsigned long long h[24], z[24] = {0};
        cudaMemcpyFromSymbol(h, g_wd_prof, sizeof h);
        cudaMemcpyToSymbol(g_wd_prof, z, sizeof z);
        static const char *names[24] = {"compact", "rect_tri", "", "far+sync", "chain", "sync", "record", "backtrace", "plain_step",
                                        "window_loop", "far:LB", "far:L1", "far:L2L3"};
        double tot = 0;
        for (int i = 0; i < 13; ++i) tot += (double)h[i];
        fprintf(stderr, "[wd_prof] nwin=%lld grid=%lld:", (long long)nwin, (long long)grid);
        for (int i = 0; i < 13; ++i) if (h[i]) fprintf(stderr, " %s=%.1f%%", names[i], 100.0 * (double)h[i] / tot);
        fprintf(stderr, " | cycles/CTA=%.3g\n", tot / (double)grid);
    }
#endif
    return PASIO_OK;
}
