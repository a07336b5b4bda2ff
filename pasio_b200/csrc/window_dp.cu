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
constexpr int PR_LIST = 64;     // pruned path: per-row survivor list capacity (entries = far block numbers)

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
    int near;                   // pruned path: columns this close to the row block are always evaluated (multiple of 32)
    u64 *cells;                 // algorithmic cells N(N-1)/2
    u64 *cells_skipped;         // cells proven irrelevant by the far-column bound (0 without pruning)
    unsigned *work_counter;
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
           + (capr / PR_FB) * 16    // sBlkI: (C first, C last, L last) of every 8-column block   (back-trace: sJump lives here)
           + (capr / PR_FB) * 8     // sBMax: max P of every 8-column block
           + 32 * 8 + 32 * 4        // sFarV, sFarA
           + 32 * PR_LIST * 2       // survivor lists, one per row of the block
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


// Far columns of one 32-row block.  Warp w owns rows 4w..4w+3.
//   lower bounds : LB_j = max( best over the near columns (sPartV),  the "split at every candidate"
//                  path through the block's own rows ) -- both are feasible segmentations; the path
//                  value is formed with a parallel prefix sum and lowered by delta to stay below the
//                  sequentially rounded value the chain would produce.
//   F1 (bounds)  : lane = (row rr, column block cc), 4 x 8 per step, UF steps in flight; surviving
//                  block numbers are appended to per-row lists in shared memory.
//   F2 (exact)   : 8-lane group g walks the list of row g: lane l8 evaluates column l8 of every
//                  surviving block, keeping its own running (max, first arg-max); the 8 lanes are
//                  merged once at the end.  Lists are flushed whenever they might overflow.
// Writes sFarV / sFarA for the warp's 4 rows; returns the number of cells skipped.
template <bool AI, int NQ>
__device__ __forceinline__ u64 far_pass(int jb, int N, int nfar, const ColRec *sCol, const int4 *sBlkI, const double *sBMax,
                                        const double *sPartV, unsigned short *sList, double *sFarV, int *sFarA,
                                        double delta, double pen,
                                        const double *__restrict__ gtab, const double *__restrict__ ltab,
                                        int alpha_int, double alpha)
{
    constexpr int UF = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, l8 = lane & 7;             // F2: row grp, column l8 of a block
    unsigned short *myList = sList + (warp * 4) * PR_LIST;

    // ---- lower bounds for all 32 rows (every warp computes them; no block-wide sync needed) ----
    double lb;
    {
        const int j = min(jb + lane, N - 1);
        const ColRec me = sCol[j], before = sCol[j - 1];
        const RowConst<AI> rj = make_row<AI>(me.C, me.L, alpha_int, alpha);
        double run = self_score<AI>(before.C, before.L, rj, gtab, ltab) + (lane ? pen : 0.0);   // w(j-1, j) [+ pen of the previous row]
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
        lb = fmax(path, nearbest);
    }
    // F1 bounds one (4 rows of this warp) x (8 columns) rectangle per lane: the counts span
    // [S_first_row - C_last_col, S_last_row - C_first_col], the shortest length is L_first_row - L_last_col,
    // and the rectangle survives unless its bound is below the SMALLEST of the 4 lower bounds.
    const int j0 = min(jb + 4 * warp, N - 1), j3 = min(jb + 4 * warp + 3, N - 1);
    const bool grp_ok = jb + 4 * warp < N;
    const ColRec mf = sCol[j0], ml = sCol[j3];
    const RowConst<AI> rcF = make_row<AI>(mf.C, mf.L, alpha_int, alpha);    // first row: smallest S, shortest lengths
    const RowConst<AI> rcL = make_row<AI>(ml.C, ml.L, alpha_int, alpha);    // last (valid) row: largest S
    double lbmin = INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double v = __shfl_sync(0xffffffffu, lb, min(4 * warp + k, 31));
        if (jb + 4 * warp + k < N) lbmin = fmin(lbmin, v);
    }
    const int r2 = 4 * warp + grp;                         // F2 row
    const bool row2_ok = jb + r2 < N;
    const ColRec me2 = sCol[min(jb + r2, N - 1)];
    const RowConst<AI> rc2 = make_row<AI>(me2.C, me2.L, alpha_int, alpha);

    // running far result of this lane (F2 role): column 0 is in no 8-block, lane l8 == 0 takes it
    double best = -INFINITY;
    int arg = 0x7fffffff;
    if (l8 == 0 && row2_ok) {
        const ColRec a = sCol[0];
        best = __dadd_rn(self_score<AI>(a.C, a.L, rc2, gtab, ltab), a.P);
        arg = 0;
    }
    u64 skipped = 0;
    int cnt = 0;                                           // list length, identical in all lanes
    const int nrow = min(4, N - (jb + 4 * warp));          // valid rows of this warp (may be <= 0)

    auto flush = [&]() {
        __syncwarp();
        constexpr int UX = 4;                              // surviving blocks in flight
        int k = 0;
        for (; k + UX <= cnt; k += UX) {
            ColRec a[UX];
            int c[UX];
            double t[UX];
#pragma unroll
            for (int u = 0; u < UX; ++u) {
                c[u] = 1 + PR_FB * (int)myList[k + u] + l8;
                a[u] = sCol[c[u]];
                t[u] = self_score<AI>(a[u].C, a[u].L, rc2, gtab, ltab);
            }
#pragma unroll
            for (int u = 0; u < UX; ++u) {
                const double tv = __dadd_rn(t[u], a[u].P);
                if (tv > best) { best = tv; arg = c[u]; }
            }
        }
        for (; k < cnt; ++k) {
            const int c0 = 1 + PR_FB * (int)myList[k] + l8;
            const ColRec a0 = sCol[c0];
            const double t0 = __dadd_rn(self_score<AI>(a0.C, a0.L, rc2, gtab, ltab), a0.P);
            if (t0 > best) { best = t0; arg = c0; }
        }
        cnt = 0;
        __syncwarp();
    };

    for (int cb0 = 0; cb0 < nfar; cb0 += 32 * UF) {
        bool surv[UF];
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            const int b = cb0 + 32 * u + lane;
            surv[u] = false;
            if (grp_ok && b < nfar) {
                // x = C of the block's first column (largest count), y = C and z = L of its last column
                const int4 blk = sBlkI[b];
                const int x_hi = rcL.cjx - blk.x, x_lo = rcF.cjx - blk.y;
                const double lg = __ldg(ltab + (rcF.lj - blk.z));
                const double s_hi = AI ? u32_to_double(x_hi) : __dsub_rn(rcL.aj, u32_to_double(blk.x));
                const double s_lo = AI ? u32_to_double(x_lo) : __dsub_rn(rcF.aj, u32_to_double(blk.y));
                const double f_hi = __dsub_rn(__ldg(gtab + x_hi), __dmul_rn(s_hi, lg));
                const double f_lo = __dsub_rn(__ldg(gtab + x_lo), __dmul_rn(s_lo, lg));
                const double ub = fmax(f_hi, f_lo) + sBMax[b] + delta;
                surv[u] = !(ub < lbmin);                            // NaN keeps the block
                if (!surv[u]) skipped += (u64)(PR_FB * nrow);
            }
        }
#pragma unroll
        for (int u = 0; u < UF; ++u) {
            const unsigned mask = __ballot_sync(0xffffffffu, surv[u]);
            if (surv[u]) myList[cnt + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)(cb0 + 32 * u + lane);
            cnt += __popc(mask);
            // the next ballot may add 32 more entries
            if (cnt > 4 * PR_LIST - 32) flush();
        }
    }
    flush();

    // merge the 8 column lanes of each row (first maximum), publish
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (l8 == 0 && row2_ok) {
        sFarV[r2] = best;
        sFarA[r2] = arg;
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
    int4 *sBlkI = reinterpret_cast<int4 *>(sPrev + capr);               // capr*2 bytes is a multiple of 16
    double *sBMax = reinterpret_cast<double *>(sBlkI + capr / PR_FB);
    double *sFarV = sBMax + capr / PR_FB;
    double *sScal = sFarV + 32;                                         // [0] magnitude of the window's largest self score, [1] max |P|
    int *sFarA = reinterpret_cast<int *>(sScal + 4);
    unsigned short *sList = reinterpret_cast<unsigned short *>(sFarA + 32);
    // the back-trace runs after the DP, when these are dead
    unsigned short *sJump = reinterpret_cast<unsigned short *>(sBlkI);  // capr*2 bytes == (capr/8)*16
    unsigned char *sMark = reinterpret_cast<unsigned char *>(sTri);     // capr <= 8192 bytes

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
        if (PRUNE) {
            for (int b = tid; PR_FB * b + PR_FB < N; b += WD_THREADS) {
                const ColRec a = sCol[1 + PR_FB * b], z = sCol[PR_FB * b + PR_FB];
                sBlkI[b] = make_int4(a.C, z.C, z.L, 0);
            }
        }
        for (int jb = 1; jb < N; jb += DP_JB) {
            constexpr int NQ = WD_WARPS / (DP_JB / (DP_RPW * RPL));
            const int near_lo = jb - p.near;
            if (!PRUNE || near_lo < 1 + PR_FB) {
                dp_block_step<AI, WD_WARPS, U, RPL>(jb, N, 0, sCol, sPrev, nullptr, sPartV, sPartA, sTri,
                                                    p.gtab, p.ltab, p.alpha_int, p.alpha, p.pen, -INFINITY, 0, 0);
            } else {
                // (1) nearest columns + triangle self scores
                block_rect_tri<AI, WD_WARPS, U, RPL>(jb, N, near_lo, sCol, sPartV, sPartA, sTri, p.gtab, p.ltab,
                                                     p.alpha_int, p.alpha);
                // (2) far columns [1, near_lo): bound blocks of 8 against a lower bound of the row maximum,
                //     evaluate the survivors exactly
                const double delta = ldexp(sScal[0] + sScal[1] + fabs(p.pen) * DP_JB, -44);
                skipped += far_pass<AI, NQ>(jb, N, (near_lo - 1) / PR_FB, sCol, sBlkI, sBMax, sPartV, sList, sFarV, sFarA,
                                            delta, p.pen, p.gtab, p.ltab, p.alpha_int, p.alpha);
                __syncthreads();
                // (3) chain: far result first (smaller columns), then the near partials, then the triangle
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
    static const int near_env = getenv("PASIO_WD_NEAR") ? atoi(getenv("PASIO_WD_NEAR")) : PR_NEAR;
    p.near = (near_env >= 32 && near_env % 32 == 0) ? near_env : PR_NEAR;
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
