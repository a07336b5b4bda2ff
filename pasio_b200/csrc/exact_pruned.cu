// K3, pruned and pipelined: the exact O(N^2) SquareSplitter DP for one long candidate list (N up to a few 1e5)
// at the speed of its dependent chain.
//
// Replaces SquareSplitter.split_without_normalizations (/root/reference/src/pasio/splitters/square_splitter.py:67-100)
// driven by all_suffixes_self_score (/root/reference/src/pasio/log_marginal_likelyhood.py:105-132).  Results (P, prev)
// are bit-identical to evaluating every cell: a cell is either evaluated in the reference's operation order or PROVED
// unable to hold the arg-max or a tie (bound.cuh).
//
// One persistent cooperative launch, one CTA per SM.  Rows in blocks of 128, steps of 32.
//   CTA 0 (diagonal) owns the dependent chain.  For row j its columns are split by distance:
//       triangle  [jb, j)           chain warp, one shuffle-broadcast step per row
//       near      [jb-32, jb)       chain warp, folded while those rows are being chained (a second accumulator)
//       mid       [F_b, jb-32)      7 helper warps, one step ahead (F_b = first row of block b - lag + 1)
//       far       [0, F_b)          worker CTAs, 'lag' - 1 blocks ahead, pruned
//     The diagonal never gathers from the tables: all self scores within distance 128 * lag of the diagonal
//     (P-independent) are produced ahead of it by workers into an L2-resident ring (S tasks) and only ADDED to the
//     finished P_i here.
//   Workers pull tasks from a fixed list: S(b, q) = a quarter of the self-score band of row block b; F(b, g) = the
//     far columns of row block b (column blocks c = g mod 8), released when block b - lag is final:
//       lower bound of every row's maximum = best exactly evaluated cell over 8 anchors (the last final row and the
//       arg-max columns of the rows before it -- the recent change points);
//       128 x 128 rectangles against the tilted corner bound, survivors again as 32 x 32, again as 4 x 8, the rest
//       evaluated exactly.  On BASELINE configs 1 and 3 under 1 % of the far cells survive (profiles/r02_*).
//   Task order makes every wait depend on tasks earlier in the list, and all CTAs are co-resident: no deadlock.
#include "bound.cuh"

#include <vector>

namespace {

constexpr int XP_RB = 128;            // rows per block
constexpr int XP_THREADS = 256;
constexpr int XP_HELP = 7;            // helper warps of the diagonal CTA
constexpr int XP_G = 8;               // far slices per row block
constexpr int XP_SQ = 4;              // self-score sub-tasks per row block
constexpr int XP_RING = 64;           // row blocks of self scores kept
constexpr int XP_SAHEAD = 16;         // self scores are produced this many blocks ahead of the diagonal
constexpr int XP_PRING = 1024;        // finished P kept in the diagonal's shared memory
constexpr int XP_ST = 63;             // self-score tile of the chain warp: distances 1..63
constexpr int XP_MB = 28;             // mid sweep: loads in flight per lane
constexpr int XP_LIST1 = 4096;

struct XpParams {
    int N, nB, nSteps, lag, DB, n_tasks, npad;
    const int32_t *L;
    const int32_t *C;
    double *P;
    int *prev;
    double *Sring;              // [ring slot][d - 1][row in block]
    int *s_ready;               // [nB] finished S sub-tasks
    int *far_ready;             // [nB] finished F slices
    int *done_block;            // blocks finished and published (P, prev, records)
    unsigned *task_counter;
    const int2 *tasks;          // x = type | block << 1, y = sub index
    double *farV;               // [XP_G][npad]
    int *farA;
    CoarseRec *rec32;           // per 32 finished columns [1 + 32q, 33 + 32q)
    CoarseRec *rec128;          // per finished block
    double *pmax;               // running max |P|
    u64 *far_cells;             // far cells evaluated exactly
    const double *gtab;
    const double *ltab;
    int alpha_int;
    double alpha, pen;
};

__device__ __forceinline__ int xp_ld_flag(const int *p) { return *reinterpret_cast<const volatile int *>(p); }

__device__ __forceinline__ void xp_wait_cta(const int *flag, int target)     // every thread of the CTA calls
{
    if (threadIdx.x == 0) {
        while (xp_ld_flag(flag) < target) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void xp_wait_helpers(const int *flag, int target)  // every helper thread (warps 1..7) calls
{
    if (threadIdx.x == 32) {
        while (xp_ld_flag(flag) < target) __nanosleep(20);
        __threadfence();
    }
    asm volatile("bar.sync 1, %0;" ::"n"(XP_HELP * 32) : "memory");
}

__device__ __forceinline__ int xp_far_bound(int b, int lag) { return b >= lag - 1 ? 1 + XP_RB * (b - lag + 1) : 0; }

__host__ __device__ inline size_t xp_diag_smem()
{
    return (size_t)XP_PRING * 8 + 3 * XP_ST * 32 * 8 + 2 * XP_HELP * 32 * 12 + 2 * XP_RB * 12 + 64;
}
__host__ __device__ inline size_t xp_worker_smem()
{
    return (size_t)XP_RB * (8 + 8 + 8 + 8)              // rows: (L, C), LB, C and L as doubles
           + 2 * XP_RB * 8                              // LB halves
           + 8 * XP_RB * 12                             // per-warp far results
           + 256 * 4 + XP_LIST1 * 4                     // level-0 / level-1 survivor lists
           + 8 * (4 + 8 + 8) + 16 * 8 + 64;             // anchors, scalars
}

// ---- diagonal CTA -----------------------------------------------------------------------------
__device__ __forceinline__ void xp_load_s_tile(const XpParams &p, int step, double *dst, int t, int nthreads)
{
    const int jb = 1 + 32 * step;
    if (jb >= p.N) return;
    const int b = step >> 2;
    const double *src = p.Sring + (size_t)(b % XP_RING) * p.DB * XP_RB + (step & 3) * 32;
    for (int c = t; c < XP_ST * 16; c += nthreads) {
        const int d1 = c >> 4, x = (c & 15) * 2;
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(src + (size_t)d1 * XP_RB + x));
        *reinterpret_cast<double2 *>(dst + d1 * 32 + x) = v;
    }
}

template <bool AI>
__device__ void xp_diagonal(const XpParams &p, unsigned char *smem)
{
    double *sP = reinterpret_cast<double *>(smem);              // [XP_PRING] finished P, ring by column index
    double *sS = sP + XP_PRING;                                 // [3][XP_ST][32] self-score tiles of three steps
    double *sMidV = sS + 3 * XP_ST * 32;                        // [2][XP_HELP][32]
    double *sFarV = sMidV + 2 * XP_HELP * 32;                   // [2][128]
    int *sMidA = reinterpret_cast<int *>(sFarV + 2 * XP_RB);    // [2][XP_HELP][32]
    int *sFarA = sMidA + 2 * XP_HELP * 32;                      // [2][128]
    double *sScal = reinterpret_cast<double *>(sFarA + 2 * XP_RB);   // [0] running max |P|
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, lag = p.lag;

    if (tid == 0) {
        sP[0] = 0.0;                         // prefix_scores[0] = 0 (square_splitter.py:72)
        __stcg(p.P, 0.0);
        __stcg(p.prev, 0);
        sScal[0] = 0.0;
    }
    for (int k = tid; k < 2 * XP_HELP * 32; k += XP_THREADS) { sMidV[k] = -INFINITY; sMidA[k] = 0; }
    xp_wait_cta(p.s_ready, XP_SQ);
    xp_load_s_tile(p, 0, sS, tid, XP_THREADS);
    xp_load_s_tile(p, 1, sS + XP_ST * 32, tid, XP_THREADS);
    __syncthreads();

    // chain warp state: accumulator of the NEXT step's rows over the columns being chained now
    double best2 = -INFINITY, pmax = 0.0;
    int arg2 = 0;
    if (warp == 0 && 1 + lane < N) {         // rows of step 0 against column 0: distance = row index
        best2 = __dadd_rn(sS[lane * 32 + lane], 0.0);
        arg2 = 0;
    }
    const double tilt_c = (double)__ldg(p.C + N - 1) + p.alpha, tilt_l = (double)__ldg(p.L + N - 1);

    for (int k = 0; k < p.nSteps; ++k) {
        const int jb = 1 + 32 * k, b = k >> 2, s = k & 3;
        if (warp == 0) {
            // ---------------- chain warp: rows [jb, jb + 32) ----------------
            const int j = jb + lane;
            double best = -INFINITY;
            int arg = 0;
            if (b >= lag - 1) {                                  // far columns (the smallest indices)
                best = sFarV[(b & 1) * XP_RB + s * 32 + lane];
                arg = sFarA[(b & 1) * XP_RB + s * 32 + lane];
            }
#pragma unroll
            for (int w = XP_HELP - 1; w >= 0; --w) {             // mid columns: helper w swept the w-th distance chunk
                const double v = sMidV[((k & 1) * XP_HELP + w) * 32 + lane];
                if (v > best) { best = v; arg = sMidA[((k & 1) * XP_HELP + w) * 32 + lane]; }
            }
            if (best2 > best) { best = best2; arg = arg2; }     // near columns
            best2 = -INFINITY;
            arg2 = 0;
            const double *tri = sS + (k % 3) * XP_ST * 32;
            const double *nxt = sS + ((k + 1) % 3) * XP_ST * 32;
            const int rows = min(32, N - jb);
            const bool valid2 = jb + 32 + lane < N;
            double mine = 0.0;
#pragma unroll 8
            for (int kk = 0; kk < rows; ++kk) {
                const double pf = __dadd_rn(best, p.pen);       // prefix_scores[j] = max + segment_creation_cost
                const double pk = __shfl_sync(0xffffffffu, pf, kk);
                if (lane == kk) mine = pf;
                if (lane > kk) {
                    const double t = __dadd_rn(tri[(lane - kk - 1) * 32 + lane], pk);
                    if (t > best) { best = t; arg = jb + kk; }
                }
                if (valid2) {
                    const double t2 = __dadd_rn(nxt[(31 + lane - kk) * 32 + lane], pk);
                    if (t2 > best2) { best2 = t2; arg2 = jb + kk; }
                }
            }
            double pm = 0.0;
            if (j < N) {
                sP[j & (XP_PRING - 1)] = mine;
                __stcg(p.P + j, mine);
                __stcg(p.prev + j, arg);
                pm = fabs(mine);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) pm = fmax(pm, __shfl_xor_sync(0xffffffffu, pm, off));
            pmax = fmax(pmax, pm);
            if (lane == 0) sScal[0] = pmax;
        } else {
            // ---------------- helper warps ----------------
            const int hw = warp - 1, ht = tid - 32;
            if (hw == XP_HELP - 1 && k > 0) {
                // records of the columns finished in step k-1 (and of the block they complete), then publish
                const int jbp = jb - 32;
                fit_column_record(__ldg(p.C + jbp + lane), __ldg(p.L + jbp + lane), sP[(jbp + lane) & (XP_PRING - 1)],
                                  p.rec32 + (k - 1), tilt_c, tilt_l);
                if (s == 0) {
                    const int c0 = 1 + XP_RB * (b - 1);
                    int cc[4], ll[4];
                    double pp[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        cc[q] = __ldg(p.C + c0 + lane + 32 * q);
                        ll[q] = __ldg(p.L + c0 + lane + 32 * q);
                        pp[q] = sP[(c0 + lane + 32 * q) & (XP_PRING - 1)];
                    }
                    fit_column_record128(cc, ll, pp, p.rec128 + (b - 1), tilt_c, tilt_l);
                    __syncwarp();
                    if (lane == 0) {
                        __stcg(p.pmax, sScal[0]);
                        __threadfence();
                        *reinterpret_cast<volatile int *>(p.done_block) = b;
                    }
                }
            }
            // mid columns of the next step's rows: [F, jb), by distance, descending (= ascending column)
            if (k + 1 < p.nSteps) {
                const int jbn = jb + 32, bn = (k + 1) >> 2;
                const int F = xp_far_bound(bn, lag);
                const int j = jbn + lane;
                const double *Sb = p.Sring + (size_t)(bn % XP_RING) * p.DB * XP_RB + ((k + 1) & 3) * 32 + lane;
                const int nd = jbn - F - 1;                      // distances 33 .. jbn + 31 - F over the warp
                const int cs = (nd + XP_HELP - 1) / XP_HELP;
                const int dlo = 33 + hw * cs, dhi = min(dlo + cs, 33 + nd);
                const int lane_lo = max(dlo, 33 + lane), lane_hi = (j < N) ? min(dhi - 1, j - F) : -1;   // this row's distances
                double best = -INFINITY;
                int arg = 0;
                for (int d = dhi - 1; d >= dlo; d -= XP_MB) {
                    double v[XP_MB];
#pragma unroll
                    for (int u = 0; u < XP_MB; ++u) {
                        const int dd = d - u;
                        v[u] = (dd >= lane_lo && dd <= lane_hi) ? __ldcg(Sb + (size_t)(dd - 1) * XP_RB) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < XP_MB; ++u) {
                        const int dd = d - u;
                        if (dd >= lane_lo && dd <= lane_hi) {
                            const double t = __dadd_rn(v[u], sP[(j - dd) & (XP_PRING - 1)]);
                            if (t > best) { best = t; arg = j - dd; }
                        }
                    }
                }
                sMidV[(((k + 1) & 1) * XP_HELP + hw) * 32 + lane] = best;
                sMidA[(((k + 1) & 1) * XP_HELP + hw) * 32 + lane] = arg;
            }
            // self-score tile of step k+2
            if (k + 2 < p.nSteps) {
                if (((k + 2) & 3) == 0) xp_wait_helpers(p.s_ready + ((k + 2) >> 2), XP_SQ);
                xp_load_s_tile(p, k + 2, sS + ((k + 2) % 3) * XP_ST * 32, ht, XP_HELP * 32);
            }
            // far results of the next block
            if (s == 3 && b + 1 < p.nB && b + 1 >= lag - 1) {
                xp_wait_helpers(p.far_ready + (b + 1), XP_G);
                if (ht < XP_RB) {
                    const int j = 1 + XP_RB * (b + 1) + ht;
                    double best = -INFINITY;
                    int arg = 0x7fffffff;
                    if (j < N) {
#pragma unroll
                        for (int g = 0; g < XP_G; ++g) {
                            const double v = __ldcg(p.farV + (size_t)g * p.npad + j);
                            const int a = __ldcg(p.farA + (size_t)g * p.npad + j);
                            if (v > best || (v == best && a < arg)) { best = v; arg = a; }
                        }
                    }
                    sFarV[((b + 1) & 1) * XP_RB + ht] = best;
                    sFarA[((b + 1) & 1) * XP_RB + ht] = arg;
                }
            }
        }
        __syncthreads();
    }
}

// ---- worker CTAs -----------------------------------------------------------------------------
// S(b, q): self scores of row block b against the columns [F_b, j), distances 1 + q*DB/4 .. (q+1)*DB/4.
template <bool AI>
__device__ void xp_s_task(const XpParams &p, int b, int q, unsigned char *smem)
{
    int2 *sLC = reinterpret_cast<int2 *>(smem);                 // (L, C) of candidates [F, r0 + nrows)
    const int tid = threadIdx.x;
    const int r0 = 1 + XP_RB * b, nrows = min(XP_RB, p.N - r0), F = xp_far_bound(b, p.lag);
    if (b >= XP_RING) xp_wait_cta(p.done_block, b - XP_RING + 1);      // the ring slot's previous block is finished
    const int cnt = r0 + nrows - F;
    for (int i = tid; i < cnt; i += XP_THREADS) sLC[i] = make_int2(__ldg(p.L + F + i), __ldg(p.C + F + i));
    __syncthreads();
    const int dq = p.DB / XP_SQ, d0 = 1 + q * dq, d1 = d0 + dq;
    const int r = tid & (XP_RB - 1), half = tid >> 7, j = r0 + r;
    if (r < nrows) {
        const int2 me = sLC[j - F];
        const RowConst<AI> row = make_row<AI>(me.y, me.x, p.alpha_int, p.alpha);
        double *Sb = p.Sring + (size_t)(b % XP_RING) * p.DB * XP_RB + r;
        const int dmax = min(d1 - 1, j - F);                    // column j - d >= F
        constexpr int U = 4;
        for (int d = d0 + half; d <= dmax; d += 2 * U) {
            double g[U], lg[U];
            int idx[U], ci[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int dd = min(d + 2 * u, dmax);
                const int2 c = sLC[j - dd - F];
                idx[u] = row.cjx - c.y;
                ci[u] = c.y;
                g[u] = __ldg(p.gtab + idx[u]);
                lg[u] = __ldg(p.ltab + (row.lj - c.x));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int dd = d + 2 * u;
                if (dd <= dmax) {
                    const double sx = AI ? u32_to_double(idx[u]) : __dsub_rn(row.aj, u32_to_double(ci[u]));
                    Sb[(size_t)(dd - 1) * XP_RB] = __dsub_rn(g[u], __dmul_rn(sx, lg[u]));
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(p.s_ready + b, 1);
    }
}

struct XpRows {
    const int2 *lc;        // (L, C)
    const double *lb, *cd, *ld;
};
// min over rows [ra, rb] of LB_r + a*C_r + b*L_r
__device__ __forceinline__ double xp_row_min(const XpRows &R, int ra, int rb, double a, double b)
{
    double m = INFINITY;
    for (int r = ra; r <= rb; ++r) m = fmin(m, R.lb[r] + (a * R.cd[r] + b * R.ld[r]));
    return m;
}

// F(b, g): far columns of row block b, column blocks c = g, g + 8, ... <= b - lag (slice 0 also column 0)
template <bool AI>
__device__ void xp_f_task(const XpParams &p, int b, int g, unsigned char *smem)
{
    int2 *sRowLC = reinterpret_cast<int2 *>(smem);                      // [128]
    double *sLB = reinterpret_cast<double *>(sRowLC + XP_RB);           // [128]
    double *sCd = sLB + XP_RB, *sLd = sCd + XP_RB;                      // [128] each
    double *sHalf = sLd + XP_RB;                                        // [2][128]
    double *sWV = sHalf + 2 * XP_RB;                                    // [8][128]
    double *sAncP = sWV + 8 * XP_RB;                                    // [8]
    double *sRed = sAncP + 8;                                           // [8]
    int *sWA = reinterpret_cast<int *>(sRed + 8);                       // [8][128]
    int2 *sAncLC = reinterpret_cast<int2 *>(sWA + 8 * XP_RB);           // [8]
    int *sList0 = reinterpret_cast<int *>(sAncLC + 8);                  // [256]
    int *sList1 = sList0 + 256;                                         // [XP_LIST1]
    int *sCnt = sList1 + XP_LIST1;                                      // [4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, lag = p.lag;
    const int r0 = 1 + XP_RB * b, nrows = min(XP_RB, N - r0);
    const int F = xp_far_bound(b, lag), ncb = b - lag + 1;              // far column blocks 0 .. ncb-1
    xp_wait_cta(p.done_block, ncb);

    const int zC = __ldg(p.C + N - 1), zL = __ldg(p.L + N - 1);
    const double scale0 = fabs(__ldg(p.gtab + zC + (AI ? p.alpha_int : 0))) + ((double)zC + p.alpha) * fabs(__ldg(p.ltab + zL)) + 1.0;
    const double pmax = __ldcg(p.pmax);
    const double delta = (scale0 + pmax + fabs(p.pen) * XP_RB) * 5.684341886080802e-14;     // 2^-44
    const double tilt_c = (double)zC + p.alpha, tilt_l = (double)zL;

    if (tid < XP_RB) {
        const int j = min(r0 + tid, N - 1);
        const int2 lc = make_int2(__ldg(p.L + j), __ldg(p.C + j));
        sRowLC[tid] = lc;
        sCd[tid] = (double)lc.y;
        sLd[tid] = (double)lc.x;
    }
    if (tid < 8) {
        // anchors: the last final row and the arg-max columns of the rows before it
        const int e = F - 1;
        int a = e;
        if (tid > 0 && e > 0) a = __ldcg(p.prev + max(e - (tid - 1), 1));
        a = min(max(a, 0), e);
        sAncLC[tid] = make_int2(__ldg(p.L + a), __ldg(p.C + a));
        sAncP[tid] = a > 0 ? __ldcg(p.P + a) : 0.0;
    }
    for (int k = tid; k < 8 * XP_RB; k += XP_THREADS) { sWV[k] = -INFINITY; sWA[k] = 0x7fffffff; }
    __syncthreads();
    {
        const int r = tid & (XP_RB - 1), h = tid >> 7;
        const int2 me = sRowLC[r];
        const RowConst<AI> row = make_row<AI>(me.y, me.x, p.alpha_int, p.alpha);
        double lb = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int2 a = sAncLC[4 * h + q];
            lb = fmax(lb, __dadd_rn(self_score<AI>(a.y, a.x, row, p.gtab, p.ltab), sAncP[4 * h + q]));
        }
        sHalf[h * XP_RB + r] = lb;
        if (g == 0 && h == 0 && r < nrows) {            // column 0 belongs to no record: always evaluated (P_0 = 0)
            sWV[r] = __dadd_rn(self_score<AI>(__ldg(p.C), __ldg(p.L), row, p.gtab, p.ltab), 0.0);
            sWA[r] = 0;
        }
    }
    __syncthreads();
    double lbabs = 0.0;
    if (tid < XP_RB) {
        double lb = fmax(sHalf[tid], sHalf[XP_RB + tid]);
        if (!(fabs(lb) < 1e300)) lb = -INFINITY;         // NaN / inf: no information for this row
        if (tid >= nrows) lb = INFINITY;                 // rows past the end never lower a minimum
        else if (lb > -INFINITY) lbabs = fabs(lb);
        sLB[tid] = lb;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lbabs = fmax(lbabs, __shfl_xor_sync(0xffffffffu, lbabs, off));
    if (lane == 0) sRed[warp] = lbabs;
    __syncthreads();
    lbabs = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) lbabs = fmax(lbabs, sRed[w]);
    const XpRows R = {sRowLC, sLB, sCd, sLd};
    const int2 rowF = sRowLC[0], rowE = sRowLC[nrows - 1];
    auto row_slack = [&](double a, double bb) { return (lbabs + fabs(a) * tilt_c + fabs(bb) * tilt_l) * 5.684341886080802e-14; };
    u64 evaluated = 0;

    for (int cbase = g; cbase < ncb; cbase += XP_THREADS * XP_G) {
        __syncthreads();
        if (tid == 0) { sCnt[0] = 0; sCnt[1] = 0; }
        __syncthreads();
        // ---- level 0: 128 rows x 128 columns, one thread per column block ----
        {
            const int c = cbase + tid * XP_G;
            if (c < ncb) {
                const CoarseRec *rec = p.rec128 + c;
                const int4 ends = __ldcg(reinterpret_cast<const int4 *>(rec));        // c_first, c_last, l_first, l_last
                const double a = __ldcg(&rec->a), bb = __ldcg(&rec->b);
                const double ub = __ldcg(&rec->mpt) + tilted_box_max<AI>(rowF.y - ends.y, rowE.y - ends.x, rowF.x - ends.w, rowE.x - ends.z,
                                                                          a, bb, p.gtab, p.ltab, p.alpha_int, p.alpha);
                const double rmin = xp_row_min(R, 0, nrows - 1, a, bb) - row_slack(a, bb);
                if (!(ub - rmin + delta < 0.0)) sList0[atomicAdd(sCnt, 1)] = c;      // NaN keeps the block
            }
        }
        __syncthreads();
        const int n0 = sCnt[0];
        // ---- level 1: 32 rows x 32 columns, one thread per rectangle ----
        for (int e = tid; e < n0 * 16; e += XP_THREADS) {
            const int c = sList0[e >> 4], rg = (e >> 2) & 3, cq = e & 3;
            if (32 * rg < nrows) {
                const int q32 = 4 * c + cq;
                const CoarseRec *rec = p.rec32 + q32;
                const int4 ends = __ldcg(reinterpret_cast<const int4 *>(rec));
                const double a = __ldcg(&rec->a), bb = __ldcg(&rec->b);
                const int ra = 32 * rg, rb = min(ra + 31, nrows - 1);
                const int2 fa = sRowLC[ra], fb = sRowLC[rb];
                const double ub = __ldcg(&rec->mpt) + tilted_box_max<AI>(fa.y - ends.y, fb.y - ends.x, fa.x - ends.w, fb.x - ends.z,
                                                                          a, bb, p.gtab, p.ltab, p.alpha_int, p.alpha);
                const double rmin = xp_row_min(R, ra, rb, a, bb) - row_slack(a, bb);
                if (!(ub - rmin + delta < 0.0)) {
                    const int slot = atomicAdd(sCnt + 1, 1);
                    if (slot < XP_LIST1) sList1[slot] = q32 * 4 + rg;
                }
            }
        }
        __syncthreads();
        const int n1 = min(sCnt[1], XP_LIST1);       // n0 <= 256 blocks x 16 = 4096: never overflows
        // ---- level 2 (4 rows x 8 columns, one lane per rectangle) and exact evaluation, one warp per 32 x 32 ----
        for (int e = warp; e < n1; e += 8) {
            const int ent = sList1[e], q32 = ent >> 2, rg = ent & 3;
            const CoarseRec *rec = p.rec32 + q32;
            const double a = __ldcg(&rec->a), bb = __ldcg(&rec->b);
            const int sub = lane & 3, ra = 32 * rg + 4 * (lane >> 2);
            bool surv = false;
            if (ra < nrows) {
                const int rb = min(ra + 3, nrows - 1);
                const int i0 = 1 + 32 * q32 + 8 * sub;
                const int2 cF = make_int2(__ldg(p.L + i0), __ldg(p.C + i0)), cL = make_int2(__ldg(p.L + i0 + 7), __ldg(p.C + i0 + 7));
                const int2 fa = sRowLC[ra], fb = sRowLC[rb];
                const double m2 = tilted_box_max<AI>(fa.y - cL.y, fb.y - cF.y, fa.x - cL.x, fb.x - cF.x, a, bb,
                                                     p.gtab, p.ltab, p.alpha_int, p.alpha);
                const double m3 = xp_row_min(R, ra, rb, a, bb) - row_slack(a, bb);
                surv = !(__ldcg(&rec->mpt8[sub]) + m2 - m3 + delta < 0.0);
            }
            unsigned mask = __ballot_sync(0xffffffffu, surv);
            evaluated += (u64)__popc(mask) * 32;
            const int er = lane >> 3, ec = lane & 7;
            while (mask) {
                const int l2 = __ffs(mask) - 1;
                mask &= mask - 1;
                const int row = 32 * rg + 4 * (l2 >> 2) + er, col = 1 + 32 * q32 + 8 * (l2 & 3) + ec;
                double t = -INFINITY;
                int ta = 0x7fffffff;
                if (row < nrows) {
                    const int2 me = sRowLC[row];
                    const RowConst<AI> rc = make_row<AI>(me.y, me.x, p.alpha_int, p.alpha);
                    t = __dadd_rn(self_score<AI>(__ldg(p.C + col), __ldg(p.L + col), rc, p.gtab, p.ltab), __ldcg(p.P + col));
                    ta = col;
                }
#pragma unroll
                for (int off = 1; off < 8; off <<= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, t, off);
                    const int oa = __shfl_xor_sync(0xffffffffu, ta, off);
                    if (ob > t || (ob == t && oa < ta)) { t = ob; ta = oa; }
                }
                if (ec == 0 && row < nrows) {
                    const double cur = sWV[warp * XP_RB + row];
                    if (t > cur || (t == cur && ta < sWA[warp * XP_RB + row])) { sWV[warp * XP_RB + row] = t; sWA[warp * XP_RB + row] = ta; }
                }
            }
        }
    }
    __syncthreads();
    if (tid < nrows) {
        double best = -INFINITY;
        int arg = 0x7fffffff;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const double v = sWV[w * XP_RB + tid];
            const int a = sWA[w * XP_RB + tid];
            if (v > best || (v == best && a < arg)) { best = v; arg = a; }
        }
        __stcg(p.farV + (size_t)g * p.npad + r0 + tid, best);
        __stcg(p.farA + (size_t)g * p.npad + r0 + tid, arg);
    }
    if (lane == 0 && evaluated) atomicAdd(p.far_cells, evaluated);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(p.far_ready + b, 1);
    }
}

template <bool AI>
__global__ void __launch_bounds__(XP_THREADS, 1)
exact_pruned_kernel(XpParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int sTask;
    if (blockIdx.x == 0) {
        xp_diagonal<AI>(p, smem);
        return;
    }
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sTask = (int)atomicAdd(p.task_counter, 1u);
        __syncthreads();
        const int t = sTask;
        if (t >= p.n_tasks) return;
        const int2 task = __ldg(p.tasks + t);
        if ((task.x & 1) == 0) xp_s_task<AI>(p, task.x >> 1, task.y, smem);
        else xp_f_task<AI>(p, task.x >> 1, task.y, smem);
    }
}

// S tasks run XP_SAHEAD blocks ahead of the F tasks; F(b) sits where the diagonal finishes block b - lag.
std::vector<int2> build_tasks(int nB, int lag)
{
    std::vector<int2> t;
    auto push_s = [&](int b) { if (b < nB) for (int q = 0; q < XP_SQ; ++q) t.push_back(make_int2(b << 1, q)); };
    auto push_f = [&](int b) { if (b >= lag - 1 && b < nB) for (int g = 0; g < XP_G; ++g) t.push_back(make_int2((b << 1) | 1, g)); };
    for (int b = 0; b < XP_SAHEAD; ++b) push_s(b);
    for (int tt = 0; tt < nB; ++tt) {
        push_f(tt + lag - 1);
        push_s(tt + XP_SAHEAD);
    }
    return t;
}

template <bool AI>
int run_exact_pruned(pasio_ctx *ctx, i64 N, int lag)
{
    XpParams p;
    p.N = (int)N;
    p.nB = (int)((N - 1 + XP_RB - 1) / XP_RB);
    p.nSteps = (int)((N - 1 + 31) / 32);
    p.lag = lag;
    p.DB = XP_RB * lag;
    p.npad = (int)((N + 127) & ~(i64)127);
    const std::vector<int2> tasks = build_tasks(p.nB, lag);
    p.n_tasks = (int)tasks.size();

    const int slots = p.nB < XP_RING ? p.nB : XP_RING;
    const size_t ring_bytes = (size_t)slots * p.DB * XP_RB * 8;
    const size_t far_bytes = (size_t)XP_G * p.npad * 12;
    const size_t rec_bytes = ((size_t)p.nSteps + p.nB + 2) * sizeof(CoarseRec);
    const size_t flag_ints = (size_t)2 * p.nB + 8;
    PASIO_TRY(pasio_reserve(ctx, ctx->xpRing, ring_bytes));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpPart, far_bytes > rec_bytes ? far_bytes : rec_bytes));
    PASIO_TRY(pasio_reserve(ctx, ctx->xpRec, rec_bytes));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpMark, flag_ints * 4 + 64 > (size_t)N ? flag_ints * 4 + 64 : (size_t)N));
    PASIO_TRY(pasio_reserve(ctx, ctx->xpTasks, tasks.size() * sizeof(int2)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->xpTasks.p, tasks.data(), tasks.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));          // the host vector goes away
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->dpMark.p, 0, flag_ints * 4 + 64, ctx->stream));

    p.L = ctx->dpL.as<int32_t>();
    p.C = ctx->dpC.as<int32_t>();
    p.P = ctx->dpP.as<double>();
    p.prev = ctx->dpPrev.as<int>();
    p.Sring = ctx->xpRing.as<double>();
    int *flags = ctx->dpMark.as<int>();
    p.pmax = reinterpret_cast<double *>(flags);                 // 8 bytes
    p.far_cells = reinterpret_cast<u64 *>(flags + 2);           // 8 bytes
    p.done_block = flags + 4;
    p.task_counter = reinterpret_cast<unsigned *>(flags + 5);
    p.s_ready = flags + 8;
    p.far_ready = p.s_ready + p.nB;
    p.tasks = ctx->xpTasks.as<int2>();
    p.farV = ctx->dpPart.as<double>();
    p.farA = reinterpret_cast<int *>(p.farV + (size_t)XP_G * p.npad);
    p.rec32 = ctx->xpRec.as<CoarseRec>();
    p.rec128 = p.rec32 + p.nSteps + 1;
    p.gtab = ctx->tab[AI ? PASIO_TAB_LGAMMA : PASIO_TAB_LGAMMA_ALPHA].as<double>();
    p.ltab = ctx->tab[PASIO_TAB_LOG].as<double>();
    p.alpha_int = (int)ctx->alpha_int;
    p.alpha = ctx->alpha;
    p.pen = ctx->pen;

    const size_t smem = xp_diag_smem() > xp_worker_smem() ? xp_diag_smem() : xp_worker_smem();
    CUDA_TRY(ctx, cudaFuncSetAttribute(exact_pruned_kernel<AI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exact_pruned_kernel<AI>, XP_THREADS, smem));
    if (per_sm < 1) return pasio_fail(ctx, PASIO_E_CUDA, "pruned exact DP kernel cannot be resident");
    int workers = ctx->sm_count - 1;                            // one CTA per SM: the diagonal has an SM to itself
    if (workers > p.n_tasks) workers = p.n_tasks;
    if (workers < 1) workers = 1;
    void *args[] = {&p};
    {
        TimingScope ts(ctx, TF_EXACT_DP);
        // cooperative launch = all CTAs co-resident, which the flag waits rely on
        CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void *)exact_pruned_kernel<AI>, dim3(1 + workers), dim3(XP_THREADS), args, smem,
                                                  ctx->stream));
    }
    // cells that were evaluated: the band the diagonal adds up + the far cells that survived the bounds
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars + 8, p.far_cells, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    i64 band = 0;
    for (int b = 0; b < p.nB; ++b) {
        const i64 F = b >= lag - 1 ? 1 + (i64)XP_RB * (b - lag + 1) : 0;
        const i64 ra = 1 + (i64)XP_RB * b, rb = ra + XP_RB < N ? ra + XP_RB : N;     // rows [ra, rb)
        band += (rb - ra) * (ra - F) + (rb - ra) * (rb - ra - 1) / 2;
    }
    ctx->last_cells = N * (N - 1) / 2;
    ctx->last_cells_skipped = ctx->last_cells - band - ctx->h_scalars[8];
    return PASIO_OK;
}

}  // namespace

int launch_exact_dp_pruned(pasio_ctx *ctx, i64 N, int lag)
{
    PASIO_TRY(pasio_reserve(ctx, ctx->dpP, (size_t)N * 8));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpPrev, (size_t)N * 4));
    return ctx->alpha_is_int ? run_exact_pruned<true>(ctx, N, lag) : run_exact_pruned<false>(ctx, N, lag);
}
