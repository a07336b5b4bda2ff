// K3, pruned and pipelined: the exact O(N^2) SquareSplitter DP for one long candidate list (N up to a few 1e5)
// at the speed of its dependent chain.
//
// Replaces SquareSplitter.split_without_normalizations (/root/reference/src/pasio/splitters/square_splitter.py:67-100)
// driven by all_suffixes_self_score (/root/reference/src/pasio/log_marginal_likelyhood.py:105-132).  Results (P, prev)
// are bit-identical to evaluating every cell: a cell is either evaluated in the reference's operation order or PROVED
// unable to hold the arg-max or a tie (bound.cuh).
//
// One persistent cooperative launch, one CTA per SM.  Rows in blocks of 128, steps of 32 (defaults: lag 5, 3 N blocks).
//   CTA 0 (diagonal) owns the dependent chain and does nothing else but add.  For row j its columns are split by distance:
//       triangle  [jb, j)              chain warp, one shuffle-broadcast step per row
//       near      [jb-32, jb)          chain warp, folded while those rows are being chained (a second accumulator)
//       mid       [band_b, jb-32)      6 sweeping warps, one step ahead (band_b = first row of block b - 1)
//       N blocks  [F_b, band_b)        worker CTAs, every cell (blocks b - lag + 1 .. b - 2), released by p_block
//       far       [0, F_b)             worker CTAs, pruned (F_b = first row of block b - lag + 1), released by done_block
//     The diagonal never gathers from the tables: the self scores of the band (P-independent) are produced ahead of it by
//     workers into an L2-resident ring (S tasks, laid out per 32-row step) and only ADDED to the finished P_i here; the chain
//     warp's tile of a step and the merged far results of a block arrive by bulk copy (cp.async.bulk + mbarrier), warp 4
//     waits for them.  The records of finished columns (tilt fits, anchors, max |P|) are made by workers, too (R tasks).
//   Workers pull tasks from a fixed list: S(b, q) = a quarter of the self-score band of row block b; R(b) = records of the
//     finished block b, then done_block; N(b, g) = 16 columns of each N block; F(b, g) = the far columns of row block b
//     (column groups q = g mod 8):
//       lower bound of every row's maximum = best exactly evaluated cell over 8 anchors (the last final row and the
//       arg-max columns of the rows before it -- the recent change points);
//       128 x 128 rectangles against the tilted corner bound, survivors again as 32 x 32, again as 4 x 8, the rest
//       evaluated exactly.  On BASELINE configs 1 and 3 about 1 % of all cells are evaluated (profiles/r02_*).
//     The task that finishes a row block's last F / N slice merges the slices into one (value, column) per row.
//   Task order makes every wait depend on tasks earlier in the list or on the diagonal, and all CTAs are co-resident: no
//   deadlock (pasio_exact_task_plan + tests/test_exact_task_plan.py check the order on a CPU).
#include "bound.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int XP_RB = 128;            // rows per block
constexpr int XP_THREADS = 256;
constexpr int XP_HELP = 7;            // helper warps of the diagonal CTA: six sweep the mid columns, one (warp 4) keeps the books
constexpr int XP_MIDW = 6;
constexpr int XP_G = 8;               // far slices per row block
constexpr int XP_NG = 8;              // slices of the column blocks next to the far region, evaluated exhaustively by workers (N tasks)
constexpr int XP_GT = XP_G + XP_NG;   // result slices per row block
constexpr int XP_SQ = 4;              // self-score sub-tasks per row block
constexpr int XP_RING = 64;           // row blocks of self scores kept
constexpr int XP_SAHEAD = 16;         // self scores are produced this many blocks ahead of the diagonal
constexpr int XP_PRING = 1024;        // finished P kept in the diagonal's shared memory
constexpr int XP_ST = 63;             // self-score tile of the chain warp: distances 1..63
#ifndef XP_MB_N
#define XP_MB_N 20
#endif
constexpr int XP_MB = XP_MB_N;             // mid sweep: loads per batch and lane; two batches are prefetched (lag 3 with N tasks: a warp's
                                      // chunk is at most 38 distances).  Was 30: the unrolled batches are most of the sweeping warps' code
constexpr int XP_LIST1 = 4096;
#ifndef XP_STREAM_LOADS
#define XP_STREAM_LOADS 1      // every self score is read once: evict-first loads keep them from displacing the tables in L1
#endif
constexpr int XP_L2CHUNK = 128;       // level-1 survivors refined per pass (<= 4096 fine rectangles listed)
constexpr size_t XP_LINEAR_BYTES = (size_t)2 << 30;

// record of 32 finished columns: the bound record plus the end points of its four 8-column sub-blocks (level 2)
struct __align__(16) XpRec32 {
    CoarseRec r;
    int sub[4][4];              // c_first, c_last, l_first, l_last
};
// published with every finished block: its last row e and the arg-max columns of rows e, e-1, ..., e-6 (the recent
// change points) -- the anchors whose exactly evaluated cells bound the maxima of the rows ahead from below
struct __align__(16) XpAnchors {
    int idx[8], L[8], C[8];
    double P[8];
};
static_assert(sizeof(XpRec32) == 144 && sizeof(XpAnchors) == 160, "record layouts");

struct XpParams {
    int N, nB, nSteps, lag, DB, n_tasks, npad;
    int nb;                     // column blocks in front of the far columns (blocks b - lag + 1 ..) evaluated by worker CTAs (N tasks), not swept by the diagonal
    int dbg;                    // PASIO_XD_DBG (timing experiments only, results become wrong): 1 no mid sweep, 2 no records,
                                // 4 no waiting for far results, 8 no tile moves, 16 mid: loads only, 32 mid: arithmetic only
    int s_slots;                // row blocks of self scores held: nB (every block has its own slab: written once per launch,
                                // so the diagonal may read it through L1) or XP_RING (slabs reused: L2-coherent loads only)
    const int32_t *L;
    const int32_t *C;
    double *P;
    int *prev;
    double *Sring;              // [ring slot][step of the block (4)][d - 1][row of the step (32)]: the tile of a step -- and every
                                // run of distances of its 32 rows -- is one contiguous piece (one bulk copy)
    int *s_ready;               // [nB] finished S sub-tasks
    int *far_ready;             // [nB] finished F and N slices
    int *done_block;            // blocks finished and published (P, prev, records)
    int *p_block;               // blocks whose P and prev are final in global memory: published a record-fit earlier than done_block (N tasks)
    unsigned *task_counter;
    const int2 *tasks;          // x = type | block << 1, y = sub index
    double *farVm;              // [nB][128] the slices merged (first maximum) by the last task of the block; farAm likewise
    int *farAm;
    int *far_done;              // [nB] 1: farVm / farAm of the block are complete
    double *farV;               // [nB][XP_GT][128]: far slices, then N slices
    int *farA;
    XpRec32 *rec32;             // per 32 finished columns [1 + 32q, 33 + 32q)
    CoarseRec *rec128;          // per finished block
    XpAnchors *anchors;         // per finished block
    double *pmax;               // running max |P|
    u64 *far_cells;             // far cells evaluated exactly
    u64 *prof;                  // [64] cycle counters; [32 + 4 * role + s]: cycles from the top of a step to the step barrier by step type
                                // s = k & 3 (role 0 chain, 1 sweeping, 2 book-keeping warp), [44] looks at the self-score flags, [48 + s] step duration by type
    const double *gtab;
    const double *ltab;
    int alpha_int;
    double alpha, pen;
    const double *consts;       // [0] scale of the bound's delta (largest |G| + s*Lg a cell can reach), [1], [2] tilt scales
};

// cycle counters kept by one thread per role (negligible cost: a few clock reads per 32-row step / per task)
enum { XQ_CHAIN = 0, XQ_CHAIN_BAR, XQ_H6_REC, XQ_H6_MID, XQ_H6_TILE, XQ_H6_SWAIT, XQ_H6_FAR, XQ_H6_BAR,
       XQ_H0_MID, XQ_H0_TILE, XQ_H0_SWAIT, XQ_H0_FAR, XQ_H0_BAR, XQ_DIAG_TOTAL,
       XQ_S_WORK = 16, XQ_S_WAIT, XQ_S_COUNT, XQ_F_WORK, XQ_F_WAIT, XQ_F_COUNT, XQ_F_MAX, XQ_F_L0, XQ_F_L1, XQ_F_L23, XQ_F_HEAD };
// PROF = false (the kernels that run unless PASIO_XD_PROF is set): no clock reads -- each is an asm volatile with a memory
// clobber that the compiler schedules nothing across; without them the kernel is 1.6 - 2.2 % faster (gpurun session k3)
template <bool PROF>
__device__ __forceinline__ long long xp_clock()
{
    if (!PROF) return 0;
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}

__device__ __forceinline__ int xp_ld_flag(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ int xp_ld_acquire(const int *p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// The merged far results of a block travel from global memory into the diagonal's shared memory as two bulk copies that
// complete on an mbarrier: no register, hence no scoreboard, is tied up while they are in flight.  (A load that is issued
// early into registers does not hide its latency here: ptxas shares the six scoreboards of a warp between loads, so the
// next wait on a shared slot also waits for the far load -- ~1.5 k cycles to L2 and back while 147 worker CTAs gather.
// Measured: profiles/r02_exact_dp_v9_notes.txt.)
__device__ __forceinline__ unsigned xp_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xp_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xp_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void xp_mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xp_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xp_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(xp_smem_addr(dst)), "l"(src), "r"(bytes), "r"(xp_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void xp_mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n .reg .pred p;\n XP_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra XP_DONE;\n bra XP_WAIT;\n XP_DONE:\n}"
                 ::"r"(xp_smem_addr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void xp_wait_cta(const int *flag, int target)     // every thread of the CTA calls
{
    if (threadIdx.x == 0) {
        while (xp_ld_flag(flag) < target) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
}
// Self scores are written by worker CTAs and read by the diagonal after an acquire.  A slab that is written only once
// per launch cannot be stale in the reader's L1 (L1 is invalidated at kernel boundaries), so ordinary loads are safe;
// ring slabs are rewritten during the launch and must be read from L2 (ld.cg).
template <bool RING, typename T>
__device__ __forceinline__ T xp_ld_s(const T *p) { return RING ? __ldcg(p) : (XP_STREAM_LOADS ? __ldcs(p) : *p); }

// Helpers: make sure the self scores of block `need` are complete.  One warp looks at the flags of 32 consecutive blocks
// at once (they are produced far ahead of the diagonal), so the round trip to L2 is paid once per ~32 blocks.
// *known: highest block index known complete (kept per thread, uniform over the helpers).
__device__ __forceinline__ void xp_wait_s_blocks(const XpParams &p, int need, int *known, int *sKnown)
{
    if (need <= *known) return;
    if ((threadIdx.x >> 5) == 4) {
        const int lane = threadIdx.x & 31;
        if ((p.dbg & 128) && lane == 0) atomicAdd(p.prof + 44, 1ull);
        unsigned ok;
        while (true) {
            const int bq = need + lane;
            const bool ready = bq < p.nB && xp_ld_flag(p.s_ready + bq) >= XP_SQ;
            ok = __ballot_sync(0xffffffffu, ready);
            if (ok & 1u) break;
            __nanosleep(20);
        }
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");          // (the slabs are read by bulk copies, too)
        if (lane == 0) *sKnown = need + __ffs(~ok | 0x80000000u) - 2 + ((ok == 0xffffffffu) ? 1 : 0);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(XP_HELP * 32) : "memory");
    *known = *sKnown;
}

__device__ __forceinline__ int xp_far_bound(int b, int lag) { return b >= lag - 1 ? 1 + XP_RB * (b - lag + 1) : 0; }
// first column of the band the diagonal sweeps itself (and of the self-score slabs): with N tasks one block later than the far bound
// N tasks of row block b cover the nb column blocks in front of the band, [F0 + ... ) with F0 = 1 + 128 (b - lag + 1) -- which is
// negative in the first blocks: there they start at column 0 (the "last column of block -1").  They exist once they cover more
// than column 0.  With them the band of row block b starts nb blocks later, so that its distances fit the slab (128 (lag - nb))
// in every block.
__host__ __device__ __forceinline__ bool xp_has_n(int b, int lag, int nb) { return nb > 0 && 1 + XP_RB * (b - lag + 1 + nb) > 1; }
__host__ __device__ __forceinline__ bool xp_has_far(int b, int lag, int nb) { return b >= lag - 1 || xp_has_n(b, lag, nb); }   // any results to fetch
__host__ __device__ __forceinline__ int xp_far_count(int b, int lag, int nb) { return (b >= lag - 1 ? XP_G : 0) + (xp_has_n(b, lag, nb) ? XP_NG : 0); }
__device__ __forceinline__ int xp_band_bound(int b, int lag, int nb) { return xp_has_n(b, lag, nb) ? 1 + XP_RB * (b - lag + 1 + nb) : xp_far_bound(b, lag); }

__host__ __device__ inline size_t xp_diag_smem()
{
    return (size_t)2 * XP_PRING * 8 + 3 * XP_ST * 32 * 8 + 2 * XP_HELP * 32 * 12 + 2 * XP_RB * 12 + 64 + XP_PRING * 4 + 16
           + 48;                                                // five mbarriers: the bulk copies of the far results and of the chain warp's tiles
}
__host__ __device__ inline size_t xp_worker_smem()
{
    return (size_t)XP_RB * (8 + 16 + 8 + 8 + 8)         // rows: (L, C), float record, LB, C and L as doubles
           + 2 * XP_RB * 8                              // LB halves
           + 8 * XP_RB * 12                             // per-warp far results
           + sizeof(XpAnchors)
           + 256 * 4 + XP_LIST1 * 4 + XP_L2CHUNK * 32 * 2   // level-0 / level-1 / fine-rectangle lists
           + 16 * 8 + 64;                               // scalars
}

// ---- diagonal CTA -----------------------------------------------------------------------------
template <bool RING>
__device__ __forceinline__ void xp_load_s_tile(const XpParams &p, int step, double *dst, int t, int nthreads)
{
    const int jb = 1 + 32 * step;
    if (jb >= p.N) return;
    const int b = step >> 2;
    const double *src = p.Sring + ((size_t)(b % p.s_slots) * 4 + (step & 3)) * p.DB * 32;
    constexpr int NV = (XP_ST * 16 + XP_HELP * 32 - 1) / (XP_HELP * 32);      // chunks per thread with 224 threads (5)
    double2 v[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) {                                            // all loads in flight, then the stores
        const int c = t + u * nthreads;
        if (c < XP_ST * 16) v[u] = xp_ld_s<RING>(reinterpret_cast<const double2 *>(src + c * 2));
    }
#pragma unroll
    for (int u = 0; u < NV; ++u) {
        const int c = t + u * nthreads;
        if (c < XP_ST * 16) *reinterpret_cast<double2 *>(dst + c * 2) = v[u];
    }
}

// Mid columns of the rows of one step, for one helper warp: the distances this warp sweeps and this lane's valid range.
struct XpMid {
    const double *Sb;      // self scores of this lane's row: Sb[(d - 1) * 32]
    int j, dlo, dhi, lane_lo, lane_hi;
};
__device__ __forceinline__ XpMid xp_mid_geometry(const XpParams &p, int step, int hw, int lane)
{
    XpMid g;
    const int jbn = 1 + 32 * step, bn = step >> 2;
    const int F = xp_band_bound(bn, p.lag, p.nb);
    g.j = jbn + lane;
    g.Sb = p.Sring + ((size_t)(bn % p.s_slots) * 4 + (step & 3)) * p.DB * 32 + lane;
    const int nd = jbn - F - 1;                          // distances 33 .. jbn + 31 - F over the warp
    const int unit = (nd + XP_MIDW - 1) / XP_MIDW;       // six equal chunks (59 distances at lag 3: two batches)
    g.dlo = 33 + hw * unit;
    g.dhi = min(g.dlo + unit, 33 + nd);
    g.lane_lo = max(g.dlo, 33 + lane);
    g.lane_hi = (g.j < p.N) ? min(g.dhi - 1, g.j - F) : -1;      // this row's distances
    return g;
}
// entries u = 0 .. XP_MB-1 of a batch are the distances d, d-1, ...; outside this row's range they hold -inf and never win.
// A batch that is valid for every lane of the warp (all but the first and the last of a sweep) is loaded without the
// per-entry range test: the sweeping warps are bound by instruction issue, not by the loads.
template <bool RING>
__device__ __forceinline__ void xp_mid_load(const XpMid &g, int d, double (&v)[XP_MB])
{
    const double *top = g.Sb + (size_t)(d - 1) * 32;                 // entry u sits u * 32 doubles below
    const int lo_all = __reduce_max_sync(0xffffffffu, g.lane_lo), hi_all = __reduce_min_sync(0xffffffffu, g.lane_hi);
    if (d - (XP_MB - 1) >= lo_all && d <= hi_all) {
#pragma unroll
        for (int u = 0; u < XP_MB; ++u) v[u] = xp_ld_s<RING>(top - u * 32);
    } else {
#pragma unroll
        for (int u = 0; u < XP_MB; ++u) {
            const int dd = d - u;
            v[u] = (dd >= g.lane_lo && dd <= g.lane_hi) ? xp_ld_s<RING>(top - u * 32) : -INFINITY;
        }
    }
}
// Five independent (max, first arg-max) chains over consecutive fifths of the batch -- written interleaved, so that the
// in-order issue finds an independent instruction every cycle -- merged in column order (a later column wins only when
// strictly greater) = the sequential first maximum.  Column of entry u: j - d + u; its P sits at a fixed offset from
// the first one in the mirrored ring.
__device__ __forceinline__ void xp_mid_fold(const XpMid &g, int d, const double (&v)[XP_MB], const double *sP, double &best, int &arg)
{
    constexpr int NCH = XP_MB % 6 == 0 ? 6 : 5, PER = XP_MB / NCH;
    static_assert(XP_MB % NCH == 0, "batch must split into equal chains");
    const int col0 = g.j - d;
    const double *pc = sP + (col0 & (XP_PRING - 1));
    double t[XP_MB];
#pragma unroll
    for (int u = 0; u < XP_MB; ++u) t[u] = __dadd_rn(v[u], pc[u]);
    double qb[NCH];
    int qu[NCH];
#pragma unroll
    for (int q = 0; q < NCH; ++q) { qb[q] = t[q * PER]; qu[q] = q * PER; }
#pragma unroll
    for (int k = 1; k < PER; ++k)
#pragma unroll
        for (int q = 0; q < NCH; ++q) {
            const bool w = t[q * PER + k] > qb[q];
            qb[q] = w ? t[q * PER + k] : qb[q];
            qu[q] = w ? q * PER + k : qu[q];
        }
#pragma unroll
    for (int q = 0; q < NCH; ++q)
        if (qb[q] > best) { best = qb[q]; arg = col0 + qu[q]; }
}

// The chain warp's self-score tile of a step (distances 1..63 of its 32 rows: 16 KB, contiguous in the slab) travels into
// shared memory as one bulk copy issued by the book-keeping warp, two steps ahead; the chain warp waits on the slot's mbarrier
// when it first reads the tile.  (Before, the six sweeping warps carried the tiles through registers: 0.5 k of their 3.4 k
// cycles per step, and they are the slowest role: profiles/r02_exact_dp_v9_notes.txt.)
__device__ __forceinline__ void xp_tile_bulk(const XpParams &p, int step, double *dst, unsigned long long *bar, int lane)
{
    const double *src = p.Sring + ((size_t)((step >> 2) % p.s_slots) * 4 + (step & 3)) * p.DB * 32;
    if (lane == 0) {                                                 // (the proxy fence sits where the slab's flag was seen: xp_wait_s_blocks)
        xp_mbar_expect(bar, XP_ST * 32 * 8);
        xp_bulk_g2s(dst, src, XP_ST * 32 * 8, bar);                  // (many small bulk copies are slow to issue: 63 of 256 bytes cost
    }                                                                //  the warp 4.5 k cycles)
}
// parity of the mbarrier phase that the bulk copy of tile `step` completes: slot step % 3; tiles 0 and 1 are loaded directly
__device__ __forceinline__ unsigned xp_tile_parity(int step) { return (unsigned)((step / 3 - (step % 3 == 2 ? 0 : 1)) & 1); }

template <bool AI, bool RING, bool PROF>
__device__ void xp_diagonal(const XpParams &p, unsigned char *smem)
{
    double *sP = reinterpret_cast<double *>(smem);              // [2 * XP_PRING] finished P, ring by column index, every entry
                                                                // mirrored XP_PRING further: a run of columns never wraps
    double *sS = sP + 2 * XP_PRING;                                 // [3][XP_ST][32] self-score tiles of three steps
    double *sMidV = sS + 3 * XP_ST * 32;                        // [2][XP_HELP][32]
    double *sFarV = sMidV + 2 * XP_HELP * 32;                   // [2][128]
    int *sMidA = reinterpret_cast<int *>(sFarV + 2 * XP_RB);    // [2][XP_HELP][32]
    int *sFarA = sMidA + 2 * XP_HELP * 32;                      // [2][128]
    double *sScal = reinterpret_cast<double *>(sFarA + 2 * XP_RB);   // [0] running max |P|
    int *sPrevRing = reinterpret_cast<int *>(sScal + 8);             // [XP_PRING] arg-max columns of the finished rows, ring by row index
    unsigned long long *sFarBar = reinterpret_cast<unsigned long long *>(sPrevRing + XP_PRING + 4);   // [2], behind sKnown (16 bytes)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, lag = p.lag;

    if (tid == 0) {
        sP[0] = 0.0;                         // prefix_scores[0] = 0 (square_splitter.py:72)
        sP[XP_PRING] = 0.0;
        __stcg(p.P, 0.0);
        __stcg(p.prev, 0);
        sScal[0] = 0.0;
#pragma unroll
        for (int q = 0; q < 5; ++q) xp_mbar_init(sFarBar + q, 1);       // [0], [1] far results; [2 + slot] tiles
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = tid; k < 2 * XP_HELP * 32; k += XP_THREADS) { sMidV[k] = -INFINITY; sMidA[k] = 0; }
    xp_wait_cta(p.s_ready, XP_SQ);
    xp_load_s_tile<RING>(p, 0, sS, tid, XP_THREADS);
    xp_load_s_tile<RING>(p, 1, sS + XP_ST * 32, tid, XP_THREADS);
    __syncthreads();

    int far_first = 0;                       // first row block with far / N results (the first use of its slot's mbarrier)
    while (far_first < p.nB && !xp_has_far(far_first, lag, p.nb)) ++far_first;
    // chain warp state: accumulator of the NEXT step's rows over the columns being chained now
    double best2 = -INFINITY;
    int arg2 = 0;
    if (warp == 0 && 1 + lane < N) {         // rows of step 0 against column 0: distance = row index
        best2 = __dadd_rn(sS[lane * 32 + lane], 0.0);
        arg2 = 0;
    }
    // helper state carried from step to step: the self scores of the NEXT mid sweep and of the next tile are loaded
    // one step ahead (they do not depend on P), so a step only adds the finished P to values that have already arrived
    const int hw_mine = warp < 4 ? warp - 1 : (warp == 4 ? XP_MIDW : warp - 2);       // 0..5 sweep, 6 = the book-keeping warp
    double pre0[XP_MB], pre1[XP_MB];
    int s_known = 0;                                     // block 0 is complete (waited above)
    int *sKnown = sPrevRing + XP_PRING;
    if (warp > 0 && hw_mine < XP_MIDW) {
        if (1 < p.nSteps) {
            const XpMid g = xp_mid_geometry(p, 1, hw_mine, lane);
            xp_mid_load<RING>(g, g.dhi - 1, pre0);
            xp_mid_load<RING>(g, g.dhi - 1 - XP_MB, pre1);
        }
    }
    long long pq[8] = {0, 0, 0, 0, 0, 0, 0, 0};       // per-thread phase cycles (only threads 0, 32 and 224 report)
    const long long t_begin = xp_clock<PROF>();
    long long t_last = t_begin, t_prev_top = t_begin;
    for (int k = 0; k < p.nSteps; ++k) {
        const int jb = 1 + 32 * k, b = k >> 2, s = k & 3;
        long long tq = xp_clock<PROF>();
        const long long t_top = tq;
        pq[6] += tq - t_last;                              // (between the clock read behind the barrier and this one)
        if (warp == 0) {
            // ---------------- chain warp: rows [jb, jb + 32) ----------------
            const int j = jb + lane;
            double best = -INFINITY;
            int arg = 0;
            if (xp_has_far(b, lag, p.nb)) {                      // far columns (the smallest indices)
                // (bulk-copied by the book-keeping warp during the previous step, which also waited for them)
                best = sFarV[(b & 1) * XP_RB + s * 32 + lane];
                arg = sFarA[(b & 1) * XP_RB + s * 32 + lane];
            }
            {   // mid columns: helper w swept the w-th distance chunk (all twelve reads first, then the merge in column order)
                double mv[XP_MIDW];
                int ma[XP_MIDW];
#pragma unroll
                for (int w = 0; w < XP_MIDW; ++w) {
                    mv[w] = sMidV[((k & 1) * XP_HELP + w) * 32 + lane];
                    ma[w] = sMidA[((k & 1) * XP_HELP + w) * 32 + lane];
                }
#pragma unroll
                for (int w = XP_MIDW - 1; w >= 0; --w)
                    if (mv[w] > best) { best = mv[w]; arg = ma[w]; }
            }
            if (best2 > best) { best = best2; arg = arg2; }     // near columns
            best2 = -INFINITY;
            arg2 = 0;
            const double *tri = sS + (k % 3) * XP_ST * 32;
            const double *nxt = sS + ((k + 1) % 3) * XP_ST * 32;
            const bool valid2 = jb + 32 + lane < N;
            double mine = 0.0;
            // Always 32 iterations: in the last (partial) step the lanes past the end only produce values nobody reads.
            // The operands of 8 iterations are fetched from shared memory ahead of the dependent chain; a lane that
            // must not take part in an update holds -inf there, so the updates need no branches.  (Rolled into 4 x 8 rows
            // the loop is 12 KB shorter and 8 cycles per row slower: profiles/r02_exact_dp_v9_notes.txt.)
            // bp = best + pen is carried along: the next row's P is (w ? t : best) + pen = w ? (t + pen) : (best + pen), bit for
            // bit, and t + pen is formed beside the compare instead of behind the select -- one DADD less on the dependent path
            double bp = __dadd_rn(best, p.pen);
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 8) {
                double tv[8], nv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int kk = k0 + u;
                    tv[u] = lane > kk ? tri[(lane - kk - 1) * 32 + lane] : -INFINITY;
                    nv[u] = valid2 ? nxt[(31 + lane - kk) * 32 + lane] : -INFINITY;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int kk = k0 + u;
                    const double pk = __shfl_sync(0xffffffffu, bp, kk);     // prefix_scores[j] = max + segment_creation_cost
                    mine = lane == kk ? bp : mine;
                    const double t = __dadd_rn(tv[u], pk);
                    const double tp = __dadd_rn(t, p.pen);
                    const bool w = t > best;
                    best = w ? t : best;
                    bp = w ? tp : bp;
                    arg = w ? jb + kk : arg;
                    const double t2 = __dadd_rn(nv[u], pk);
                    const bool w2 = t2 > best2;
                    best2 = w2 ? t2 : best2;
                    arg2 = w2 ? jb + kk : arg2;
                }
            }
            if (j < N) {
                sP[j & (XP_PRING - 1)] = mine;
                sP[(j & (XP_PRING - 1)) + XP_PRING] = mine;
                __stcg(p.P + j, mine);
                __stcg(p.prev + j, arg);
            }
            { const long long t1 = xp_clock<PROF>(); pq[0] += t1 - tq; tq = t1; }
        } else {
            // ---------------- helper warps ----------------
            // warps 1,2,3,5,6,7 sweep the mid columns (six distance chunks, in this order) and move the tiles; warp 4 -- which
            // shares its scheduler with the chain warp -- keeps the books: records, anchors, publishing, the far results
            const int hw = hw_mine;
            if (hw == XP_MIDW) {
                // (first: the sweeping warps wait at a named barrier for this warp's look at the flags)
                if (k + 3 < p.nSteps && ((k + 3) & 3) == 0) xp_wait_s_blocks(p, (k + 3) >> 2, &s_known, sKnown);
                { const long long t1 = xp_clock<PROF>(); pq[3] += t1 - tq; tq = t1; }
                // P and prev of the block that finished with the previous step are in global memory (the chain warp stored them
                // before the step barrier): the N tasks need nothing else, so they are released before the records are made
                if (s == 0 && k > 0 && lane == 0) {
                    __threadfence();
                    *reinterpret_cast<volatile int *>(p.p_block) = b;
                }
                // The far results of the next block: the last worker task of the block has merged the slices into one (value,
                // column) per row.  This warp waits for that and starts two bulk copies into the diagonal's shared memory; the
                // chain warp waits for them on the mbarrier when it starts the block, a step later.  First thing in the step:
                // the copies have the whole step, and nobody else does anything for the far results any more (the sweeping
                // warps fetched and merged 16 slices per row: 2.6 k cycles per block on the critical path,
                // profiles/r02_exact_dp_v9_notes.txt).
                if (s == 3 && b + 1 < p.nB && xp_has_far(b + 1, lag, p.nb)) {
                    if (lane == 0) {
                        if (!(p.dbg & 4)) while (xp_ld_acquire(p.far_done + (b + 1)) == 0) __nanosleep(20);
                        asm volatile("fence.proxy.async;" ::: "memory");
                        unsigned long long *bar = sFarBar + ((b + 1) & 1);
                        xp_mbar_expect(bar, XP_RB * 12);
                        xp_bulk_g2s(sFarV + ((b + 1) & 1) * XP_RB, p.farVm + (size_t)(b + 1) * XP_RB, XP_RB * 8, bar);
                        xp_bulk_g2s(sFarA + ((b + 1) & 1) * XP_RB, p.farAm + (size_t)(b + 1) * XP_RB, XP_RB * 4, bar);
                    }
                    __syncwarp();
                }
                // ... and waits for the copies it started (a slot's n-th use completes phase n of its mbarrier), so that the step
                // barrier hands the far results to the chain warp
                if (s == 3 && b + 1 < p.nB && xp_has_far(b + 1, lag, p.nb)) xp_mbar_wait(sFarBar + ((b + 1) & 1), (unsigned)(((b + 1 - far_first) >> 1) & 1));
                // the chain warp's tile of the step after next, started by the first sweeping warp at the top of this step: this
                // warp has the time to wait for it, so that the step barrier hands it to the chain warp
                if (k + 2 < p.nSteps && !(p.dbg & 8)) xp_mbar_wait(sFarBar + 2 + (k + 2) % 3, xp_tile_parity(k + 2));
                { const long long t1 = xp_clock<PROF>(); pq[4] += t1 - tq; tq = t1; }
                // (the records of the finished columns, the anchors and done_block are made by worker CTAs: R tasks)
                { const long long t1 = xp_clock<PROF>(); pq[0] += t1 - tq; tq = t1; }
            } else {
                // self scores of block (k+3)/4 must be complete before anything of step k+3 is prefetched below
                if (k + 3 < p.nSteps && ((k + 3) & 3) == 0) xp_wait_s_blocks(p, (k + 3) >> 2, &s_known, sKnown);
                { const long long t1 = xp_clock<PROF>(); pq[3] += t1 - tq; tq = t1; }
                // the chain warp's tile of the step after next into the slot it stopped reading a step ago: started by the first
                // sweeping warp; the book-keeping warp waits for it before the step barrier
                const bool tile_mine = hw == 0 && k + 2 < p.nSteps && !(p.dbg & 8);
                if (tile_mine) xp_tile_bulk(p, k + 2, sS + ((k + 2) % 3) * XP_ST * 32, sFarBar + 2 + (k + 2) % 3, lane);
                // mid columns of the next step's rows: [F, jb), by distance, descending (= ascending column).  The two
                // batches were loaded during the previous step; as soon as a batch is folded its registers take the loads
                // of the step after, so the transfers run under the arithmetic
                const bool have1 = k + 1 < p.nSteps && !(p.dbg & 1), have2 = k + 2 < p.nSteps && !(p.dbg & 1);
                const XpMid g1 = xp_mid_geometry(p, have1 ? k + 1 : k, hw, lane);
                const XpMid g2 = xp_mid_geometry(p, have2 ? k + 2 : k, hw, lane);
                double best = -INFINITY;
                int arg = 0;
                const bool do_fold = !(p.dbg & 16), do_load = !(p.dbg & 32);
                if (have1 && do_fold) xp_mid_fold(g1, g1.dhi - 1, pre0, sP, best, arg);
                if (have2 && do_load) xp_mid_load<RING>(g2, g2.dhi - 1, pre0);
                // (with N tasks a warp's chunk is 22 .. 38 distances: the second batch is empty in half of the steps)
                const bool two1 = g1.dhi - 1 - XP_MB >= g1.dlo, two2 = g2.dhi - 1 - XP_MB >= g2.dlo;      // warp-uniform
                if (have1) {
                    if (do_fold && two1) xp_mid_fold(g1, g1.dhi - 1 - XP_MB, pre1, sP, best, arg);
                    for (int d = g1.dhi - 1 - 2 * XP_MB; d >= g1.dlo; d -= XP_MB) {    // (lag 4: further batches, loaded here)
                        double v[XP_MB];
                        xp_mid_load<RING>(g1, d, v);
                        xp_mid_fold(g1, d, v, sP, best, arg);
                    }
                    sMidV[(((k + 1) & 1) * XP_HELP + hw) * 32 + lane] = best;
                    sMidA[(((k + 1) & 1) * XP_HELP + hw) * 32 + lane] = arg;
                }
                if (have2 && do_load && two2) xp_mid_load<RING>(g2, g2.dhi - 1 - XP_MB, pre1);
                if (p.dbg & 16) { double acc = 0.0; for (int u = 0; u < XP_MB; ++u) acc += pre0[u] + pre1[u]; if (acc == 1.2345) sMidV[0] = acc; }   // (keeps the loads alive)
                { const long long t1 = xp_clock<PROF>(); pq[1] += t1 - tq; tq = t1; }
                { const long long t1 = xp_clock<PROF>(); pq[2] += t1 - tq; tq = t1; }
            }
        }
        if ((p.dbg & 128) && lane == 0 && (warp == 0 || warp == 1 || warp == 4)) {
            const long long ta = xp_clock<PROF>();
            atomicAdd(p.prof + 32 + 4 * (warp == 0 ? 0 : (warp == 1 ? 1 : 2)) + s, (u64)(ta - t_top));
            if (warp == 0 && k > 0) atomicAdd(p.prof + 48 + ((k - 1) & 3), (u64)(t_top - t_prev_top));
        }
        t_prev_top = t_top;
        __syncthreads();
        t_last = xp_clock<PROF>();
        pq[5] += t_last - tq;
    }
    if ((p.dbg & 64) && lane == 0) printf("[xp_dbg] warp %d: barrier gap %lld, phases %lld %lld %lld %lld %lld %lld cycles/step\n", warp, pq[6] / p.nSteps, pq[0] / p.nSteps, pq[1] / p.nSteps, pq[2] / p.nSteps, pq[3] / p.nSteps, pq[4] / p.nSteps, pq[5] / p.nSteps);
    if (tid == 0) { p.prof[XQ_CHAIN] = pq[0]; p.prof[XQ_CHAIN_BAR] = pq[5]; p.prof[XQ_DIAG_TOTAL] = xp_clock<PROF>() - t_begin; p.prof[31] = pq[6]; }
    if (tid == 32) p.prof[XQ_H6_MID] = pq[6];                // (the sweeping warp's back-edge gap, in a slot the book-keeping warp never uses)
    if (tid == 32) { p.prof[XQ_H0_MID] = pq[1]; p.prof[XQ_H0_TILE] = pq[2]; p.prof[XQ_H0_SWAIT] = pq[3]; p.prof[XQ_H0_FAR] = pq[4]; p.prof[XQ_H0_BAR] = pq[5]; }
    if (warp > 0 && hw_mine < XP_MIDW && lane == 0) p.prof[hw_mine < 2 ? 14 + hw_mine : 25 + hw_mine] = pq[1];      // mid cycles of every sweeping warp (slots 14, 15, 27 .. 30)
    if (tid == 128) { p.prof[XQ_H6_REC] = pq[0]; p.prof[XQ_H6_TILE] = pq[6]; /* (its back-edge gap; it never sweeps or moves tiles) */ p.prof[XQ_H6_SWAIT] = pq[3]; p.prof[XQ_H6_FAR] = pq[4]; p.prof[XQ_H6_BAR] = pq[5]; }
}

// ---- worker CTAs -----------------------------------------------------------------------------
// S(b, q): self scores of row block b against the columns [F_b, j), distances 1 + q*DB/4 .. (q+1)*DB/4.
template <bool AI, bool PROF>
__device__ void xp_s_task(const XpParams &p, int b, int q, unsigned char *smem)
{
    int2 *sLC = reinterpret_cast<int2 *>(smem);                 // (L, C) of candidates [F, r0 + nrows)
    const int tid = threadIdx.x;
    const int r0 = 1 + XP_RB * b, nrows = min(XP_RB, p.N - r0), F = xp_band_bound(b, p.lag, p.nb);
    const long long ts0 = xp_clock<PROF>();
    if (b >= p.s_slots) xp_wait_cta(p.done_block, b - p.s_slots + 1);  // ring only: the slot's previous block is finished
    const long long ts1 = xp_clock<PROF>();
    const int cnt = r0 + nrows - F;
    for (int i = tid; i < cnt; i += XP_THREADS) sLC[i] = make_int2(__ldg(p.L + F + i), __ldg(p.C + F + i));
    __syncthreads();
    const int dq = p.DB / XP_SQ, d0 = 1 + q * dq, d1 = d0 + dq;
    const int r = tid & (XP_RB - 1), half = tid >> 7, j = r0 + r;
    if (r < nrows) {
        const int2 me = sLC[j - F];
        const RowConst<AI> row = make_row<AI>(me.y, me.x, p.alpha_int, p.alpha);
        double *Sb = p.Sring + ((size_t)(b % p.s_slots) * 4 + (r >> 5)) * p.DB * 32 + (r & 31);
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
                    Sb[(dd - 1) * 32] = __dsub_rn(g[u], __dmul_rn(sx, lg[u]));
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(p.s_ready + b, 1);
        atomicAdd(p.prof + XQ_S_WAIT, (u64)(ts1 - ts0));
        atomicAdd(p.prof + XQ_S_WORK, (u64)(xp_clock<PROF>() - ts1));
        atomicAdd(p.prof + XQ_S_COUNT, 1ull);
    }
}

struct XpRows {
    const int2 *lc;        // (L, C)
    const double *lb, *cd, *ld;
};
// min over rows [ra, rb] of LB_r + a*C_r + b*L_r
__device__ __forceinline__ double xp_row_min(const XpRows &R, int ra, int rb, double a, double b)
{
    double m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;      // four chains: the minimum is order-free
    int r = ra;
    for (; r + 3 <= rb; r += 4) {
        m0 = fmin(m0, R.lb[r] + (a * R.cd[r] + b * R.ld[r]));
        m1 = fmin(m1, R.lb[r + 1] + (a * R.cd[r + 1] + b * R.ld[r + 1]));
        m2 = fmin(m2, R.lb[r + 2] + (a * R.cd[r + 2] + b * R.ld[r + 2]));
        m3 = fmin(m3, R.lb[r + 3] + (a * R.cd[r + 3] + b * R.ld[r + 3]));
    }
    for (; r <= rb; ++r) m0 = fmin(m0, R.lb[r] + (a * R.cd[r] + b * R.ld[r]));
    return fmin(fmin(m0, m1), fmin(m2, m3));
}

// Every F and N task of a row block ends here: the task that finishes last merges the block's slices -- first maximum: the
// larger value, on a tie the smaller column -- into one (value, column) per row, farVm / farAm, and raises far_done.  The
// diagonal copies those 1.5 KB into its shared memory in bulk; before, its sweeping warps fetched and merged the 16 slices.
__device__ void xp_far_finish(const XpParams &p, int b)
{
    __shared__ int sLast;
    const int tid = threadIdx.x;
    __syncthreads();                                                    // this task's slice is written
    if (tid == 0) {
        int before;                                                     // release of this slice + acquire of the others' in one
        asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(before) : "l"(p.far_ready + b) : "memory");
        sLast = before == xp_far_count(b, p.lag, p.nb) - 1;
    }
    __syncthreads();
    if (!sLast) return;
    if (tid < XP_RB) {
        double best = -INFINITY;
        int arg = 0x7fffffff;
        if (1 + XP_RB * b + tid < p.N) {
            double v[XP_GT];
            int a[XP_GT];
#pragma unroll
            for (int g = 0; g < XP_GT; ++g) {
                const bool have = g < XP_G ? b >= p.lag - 1 : xp_has_n(b, p.lag, p.nb);
                v[g] = have ? __ldcg(p.farV + ((size_t)b * XP_GT + g) * XP_RB + tid) : -INFINITY;
                a[g] = have ? __ldcg(p.farA + ((size_t)b * XP_GT + g) * XP_RB + tid) : 0x7fffffff;
            }
#pragma unroll
            for (int g = 0; g < XP_GT; ++g)
                if (v[g] > best || (v[g] == best && a[g] < arg)) { best = v[g]; arg = a[g]; }
        }
        __stcg(p.farVm + (size_t)b * XP_RB + tid, best);
        __stcg(p.farAm + (size_t)b * XP_RB + tid, arg);
    }
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], 1;" ::"l"(p.far_done + b) : "memory");
}

// F(b, g): far columns of row block b that lie in the 32-column groups q = g, g + 8, ... (slice 0 also column 0): the
// four groups of a column block belong to four slices, so the blocks next to the band -- where most survivors are --
// are shared out evenly.
// Written for LATENCY (the diagonal reaches block b two or three blocks after this task is released): every level
// first issues all the loads of all its rectangles, then consumes them, so a level costs one or two round trips to
// L2 whatever its size.
template <bool AI, bool PROF>
__device__ void xp_f_task(const XpParams &p, int b, int g, unsigned char *smem)
{
    int2 *sRowLC = reinterpret_cast<int2 *>(smem);                      // [128]
    float4 *sRowF = reinterpret_cast<float4 *>(sRowLC + XP_RB);         // [128] (LB, C, L) as floats: level-0 minima
    double *sLB = reinterpret_cast<double *>(sRowF + XP_RB);            // [128]
    double *sCd = sLB + XP_RB, *sLd = sCd + XP_RB;                      // [128] each
    double *sHalf = sLd + XP_RB;                                        // [2][128]
    double *sWV = sHalf + 2 * XP_RB;                                    // [8][128]
    double *sRed = sWV + 8 * XP_RB;                                     // [8]
    XpAnchors *sAnc = reinterpret_cast<XpAnchors *>(sRed + 8);
    int *sWA = reinterpret_cast<int *>(sAnc + 1);                       // [8][128]
    int *sList0 = sWA + 8 * XP_RB;                                      // [256]
    int *sList1 = sList0 + 256;                                         // [XP_LIST1]
    int *sCnt = sList1 + XP_LIST1;                                      // [4]
    unsigned short *sList2 = reinterpret_cast<unsigned short *>(sCnt + 4);   // [XP_L2CHUNK * 32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, lag = p.lag;
    const int r0 = 1 + XP_RB * b, nrows = min(XP_RB, N - r0);
    const int ncb = b - lag + 1;                                        // far column blocks 0 .. ncb-1
    const long long tf0 = xp_clock<PROF>();
    xp_wait_cta(p.done_block, ncb);
    const long long tf1 = xp_clock<PROF>();
    long long tl0 = 0, tl1 = 0, tl2 = 0, thead = 0;

    // ---- phase A: rows, anchors, this thread's first level-0 record, scalars: one round trip ----
    const double scale0 = __ldg(p.consts), tilt_c = __ldg(p.consts + 1), tilt_l = __ldg(p.consts + 2);
    const double pmax = __ldcg(p.pmax);
    int2 rowlc = make_int2(0, 0);
    if (tid < XP_RB) {
        const int j = min(r0 + tid, N - 1);
        rowlc = make_int2(__ldg(p.L + j), __ldg(p.C + j));
    }
    if (ncb >= 1) {
        if (tid < (int)(sizeof(XpAnchors) / 4))
            reinterpret_cast<int *>(sAnc)[tid] = __ldcg(reinterpret_cast<const int *>(p.anchors + (ncb - 1)) + tid);
    } else if (tid < (int)(sizeof(XpAnchors) / 4)) {
        reinterpret_cast<int *>(sAnc)[tid] = 0;                         // only column 0 is behind this block: P_0 = 0
    }
    int4 ends0 = make_int4(0, 0, 0, 0);
    double a0 = 0.0, b0 = 0.0, mpt0 = 0.0;
    static_assert(XP_G == 8, "slice g owns the 32-column groups q = g (mod 8): group g & 3 of every block c = g >> 2 (mod 2)");
    const int c_first_chunk = (g >> 2) + 2 * tid;
    if (c_first_chunk < ncb) {
        const CoarseRec *rec = p.rec128 + c_first_chunk;
        ends0 = __ldcg(reinterpret_cast<const int4 *>(rec));            // c_first, c_last, l_first, l_last
        a0 = __ldcg(&rec->a);
        b0 = __ldcg(&rec->b);
        mpt0 = __ldcg(&rec->mpt);
    }
    for (int k = tid; k < 8 * XP_RB; k += XP_THREADS) { sWV[k] = -INFINITY; sWA[k] = 0x7fffffff; }
    if (tid < XP_RB) {
        sRowLC[tid] = rowlc;
        sCd[tid] = (double)rowlc.y;
        sLd[tid] = (double)rowlc.x;
    }
    const double delta = (scale0 + pmax + fabs(p.pen) * XP_RB) * 5.684341886080802e-14;     // 2^-44
    __syncthreads();

    // ---- phase B: gathers of the lower bounds (rows x anchors), of column 0 and of this thread's level-0 box ----
    const int2 rowF = sRowLC[0], rowE = sRowLC[nrows - 1];
    BoxCorners box0 = {0.0, 0.0, 0.0, 0.0};
    if (c_first_chunk < ncb)
        box0 = box_corner_loads<AI>(rowF.y - ends0.y, rowE.y - ends0.x, rowF.x - ends0.w, rowE.x - ends0.z, p.gtab, p.ltab, p.alpha_int);
    {
        const int r = tid & (XP_RB - 1), h = tid >> 7;
        const int2 me = sRowLC[r];
        const RowConst<AI> row = make_row<AI>(me.y, me.x, p.alpha_int, p.alpha);
        double gq[4], lq[4], g00 = 0.0, l00 = 0.0;
        int ci[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ci[q] = sAnc->C[4 * h + q];
            gq[q] = __ldg(p.gtab + (row.cjx - ci[q]));
            lq[q] = __ldg(p.ltab + (row.lj - sAnc->L[4 * h + q]));
        }
        const bool col0 = g == 0 && h == 0 && r < nrows;     // column 0 belongs to no record: always evaluated (P_0 = 0)
        const int c00 = __ldg(p.C), l00i = __ldg(p.L);
        if (col0) {
            g00 = __ldg(p.gtab + (row.cjx - c00));
            l00 = __ldg(p.ltab + (row.lj - l00i));
        }
        double lb = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double sx = AI ? u32_to_double(row.cjx - ci[q]) : __dsub_rn(row.aj, u32_to_double(ci[q]));
            lb = fmax(lb, __dadd_rn(__dsub_rn(gq[q], __dmul_rn(sx, lq[q])), sAnc->P[4 * h + q]));
        }
        sHalf[h * XP_RB + r] = lb;
        if (col0) {
            const double sx = AI ? u32_to_double(row.cjx - c00) : __dsub_rn(row.aj, u32_to_double(c00));
            sWV[r] = __dadd_rn(__dsub_rn(g00, __dmul_rn(sx, l00)), 0.0);
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
        // float copy for the coarsest level: rounded DOWN so that it stays a lower bound of the double value
        float lf = (float)lb;
        if ((double)lf > lb) lf = nextafterf(lf, -INFINITY);
        sRowF[tid] = make_float4(lf, (float)sRowLC[tid].y, (float)sRowLC[tid].x, 0.f);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lbabs = fmax(lbabs, __shfl_xor_sync(0xffffffffu, lbabs, off));
    if (lane == 0) sRed[warp] = lbabs;
    if (tid == 0) { sCnt[0] = 0; sCnt[1] = 0; }
    __syncthreads();
    lbabs = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) lbabs = fmax(lbabs, sRed[w]);
    const XpRows R = {sRowLC, sLB, sCd, sLd};
    auto row_slack = [&](double a, double bb) { return (lbabs + fabs(a) * tilt_c + fabs(bb) * tilt_l) * 5.684341886080802e-14; };   // 2^-44
    u64 evaluated = 0;
    thead = xp_clock<PROF>() - tf1;

    for (int cbase = g >> 2; cbase < ncb; cbase += XP_THREADS * 2) {
        long long tq = xp_clock<PROF>();
        // ---- level 0: 128 rows x 128 columns, one thread per column block; row minima in float ----
        {
            const int c = cbase + tid * 2;
            if (c < ncb) {
                int4 ends = ends0;
                double a = a0, bb = b0, mpt = mpt0;
                BoxCorners box = box0;
                if (cbase != (g >> 2)) {                 // (beyond 512 column blocks: later chunks load here)
                    const CoarseRec *rec = p.rec128 + c;
                    ends = __ldcg(reinterpret_cast<const int4 *>(rec));
                    a = __ldcg(&rec->a);
                    bb = __ldcg(&rec->b);
                    mpt = __ldcg(&rec->mpt);
                    box = box_corner_loads<AI>(rowF.y - ends.y, rowE.y - ends.x, rowF.x - ends.w, rowE.x - ends.z, p.gtab, p.ltab, p.alpha_int);
                }
                const double ub = mpt + tilted_box_max_of(box, rowF.y - ends.y, rowE.y - ends.x, rowF.x - ends.w, rowE.x - ends.z, a, bb, p.alpha);
                const float af = (float)a, bf = (float)bb;
                float m0 = INFINITY, m1 = INFINITY, m2 = INFINITY, m3 = INFINITY;
                for (int r = 0; r < XP_RB; r += 4) {     // rows past the end hold +inf
                    const float4 q0 = sRowF[r], q1 = sRowF[r + 1], q2 = sRowF[r + 2], q3 = sRowF[r + 3];
                    m0 = fminf(m0, fmaf(af, q0.y, fmaf(bf, q0.z, q0.x)));
                    m1 = fminf(m1, fmaf(af, q1.y, fmaf(bf, q1.z, q1.x)));
                    m2 = fminf(m2, fmaf(af, q2.y, fmaf(bf, q2.z, q2.x)));
                    m3 = fminf(m3, fmaf(af, q3.y, fmaf(bf, q3.z, q3.x)));
                }
                // float error: conversions of a, b, C, L and two fused roundings, each <= 2^-24 of the magnitudes involved
                const double err = (lbabs + fabs(a) * tilt_c + fabs(bb) * tilt_l) * 9.5367431640625e-07;      // 2^-20
                const double rmin = (double)fminf(fminf(m0, m1), fminf(m2, m3)) - err;
                if (!(ub - rmin + delta < 0.0)) sList0[atomicAdd(sCnt, 1)] = c;      // NaN keeps the block
            }
        }
        __syncthreads();
        { const long long t1 = xp_clock<PROF>(); tl0 += t1 - tq; tq = t1; }
        const int n0 = sCnt[0];
        // ---- level 1: 32 rows x 32 columns (this slice's column group of every surviving block), one thread per
        //      rectangle, two rectangles in flight per thread ----
        for (int e0 = tid; e0 < n0 * 4; e0 += 2 * XP_THREADS) {
            int4 ends[2];
            double a[2], bb[2], mpt[2];
            BoxCorners box[2];
            int ra[2], rb[2], ent[2];
            bool act[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = e0 + u * XP_THREADS;
                act[u] = false;
                ent[u] = 0;
                if (e < n0 * 4) {
                    const int c = sList0[e >> 2], rg = e & 3, cq = g & 3;
                    if (32 * rg < nrows) {
                        act[u] = true;
                        ent[u] = (4 * c + cq) * 4 + rg;
                        ra[u] = 32 * rg;
                        rb[u] = min(ra[u] + 31, nrows - 1);
                        const CoarseRec *rec = &p.rec32[4 * c + cq].r;
                        ends[u] = __ldcg(reinterpret_cast<const int4 *>(rec));
                        a[u] = __ldcg(&rec->a);
                        bb[u] = __ldcg(&rec->b);
                        mpt[u] = __ldcg(&rec->mpt);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (act[u]) {
                    const int2 fa = sRowLC[ra[u]], fb = sRowLC[rb[u]];
                    box[u] = box_corner_loads<AI>(fa.y - ends[u].y, fb.y - ends[u].x, fa.x - ends[u].w, fb.x - ends[u].z, p.gtab, p.ltab, p.alpha_int);
                }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (act[u]) {
                    const int2 fa = sRowLC[ra[u]], fb = sRowLC[rb[u]];
                    const double ub = mpt[u] + tilted_box_max_of(box[u], fa.y - ends[u].y, fb.y - ends[u].x, fa.x - ends[u].w, fb.x - ends[u].z,
                                                                 a[u], bb[u], p.alpha);
                    const double rmin = xp_row_min(R, ra[u], rb[u], a[u], bb[u]) - row_slack(a[u], bb[u]);
                    if (!(ub - rmin + delta < 0.0)) {
                        const int slot = atomicAdd(sCnt + 1, 1);
                        if (slot < XP_LIST1) sList1[slot] = ent[u];
                    }
                }
        }
        __syncthreads();
        { const long long t1 = xp_clock<PROF>(); tl1 += t1 - tq; tq = t1; }
        const int n1 = min(sCnt[1], XP_LIST1);       // n0 <= 256 blocks x 16 = 4096: never overflows
        // ---- level 2 (4 rows x 8 columns, one lane per rectangle) then exact evaluation, XP_L2CHUNK level-1 survivors per pass ----
        for (int base = 0; base < n1; base += XP_L2CHUNK) {
            if (tid == 0) sCnt[2] = 0;
            __syncthreads();
            const int nhere = min(XP_L2CHUNK, n1 - base);
            constexpr int UE = 4;                        // level-1 survivors in flight per warp
            for (int i0 = warp * UE; i0 < nhere; i0 += 8 * UE) {
                double a[UE], bb[UE], m8[UE];
                int4 sub[UE];
                BoxCorners box[UE];
                int ra[UE], rb[UE];
                bool act[UE];
#pragma unroll
                for (int u = 0; u < UE; ++u) {
                    act[u] = false;
                    if (i0 + u < nhere) {
                        const int ent = sList1[base + i0 + u], q32 = ent >> 2, rg = ent & 3;
                        ra[u] = 32 * rg + 4 * (lane >> 2);
                        if (ra[u] < nrows) {
                            act[u] = true;
                            rb[u] = min(ra[u] + 3, nrows - 1);
                            const XpRec32 *rec = p.rec32 + q32;
                            a[u] = __ldcg(&rec->r.a);
                            bb[u] = __ldcg(&rec->r.b);
                            m8[u] = __ldcg(&rec->r.mpt8[lane & 3]);
                            sub[u] = __ldcg(reinterpret_cast<const int4 *>(rec->sub[lane & 3]));      // c_first, c_last, l_first, l_last
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UE; ++u)
                    if (act[u]) {
                        const int2 fa = sRowLC[ra[u]], fb = sRowLC[rb[u]];
                        box[u] = box_corner_loads<AI>(fa.y - sub[u].y, fb.y - sub[u].x, fa.x - sub[u].w, fb.x - sub[u].z, p.gtab, p.ltab, p.alpha_int);
                    }
#pragma unroll
                for (int u = 0; u < UE; ++u) {
                    bool surv = false;
                    if (act[u]) {
                        const int2 fa = sRowLC[ra[u]], fb = sRowLC[rb[u]];
                        const double m2 = tilted_box_max_of(box[u], fa.y - sub[u].y, fb.y - sub[u].x, fa.x - sub[u].w, fb.x - sub[u].z, a[u], bb[u], p.alpha);
                        const double m3 = xp_row_min(R, ra[u], rb[u], a[u], bb[u]) - row_slack(a[u], bb[u]);
                        surv = !(m8[u] + m2 - m3 + delta < 0.0);
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, surv);
                    if (mask) {
                        int slot = 0;
                        if (lane == 0) slot = atomicAdd(sCnt + 2, __popc(mask));
                        slot = __shfl_sync(0xffffffffu, slot, 0);
                        if (surv) sList2[slot + __popc(mask & ((1u << lane) - 1u))] = (unsigned short)(((i0 + u) << 5) | lane);
                    }
                }
            }
            __syncthreads();
            // exact evaluation of the listed 4 x 8 rectangles, eight per warp pass: four lanes per rectangle, a lane owns one
            // row and walks its eight columns in ascending order (strict '>': the first maximum), all sixteen table
            // gathers of a lane in flight together
            const int n2 = sCnt[2];
            evaluated += (warp == 0 && lane == 0) ? (u64)n2 * 32 : 0;
            for (int f0 = warp * 8; f0 < n2; f0 += 64) {
                const int f = f0 + (lane >> 2), er = lane & 3;
                bool act = f < n2;
                int row = 0, c0 = 0;
                if (act) {
                    const int code = sList2[f], l2 = code & 31;
                    const int ent = sList1[base + (code >> 5)], q32 = ent >> 2, rg = ent & 3;
                    row = 32 * rg + 4 * (l2 >> 2) + er;
                    c0 = 1 + 32 * q32 + 8 * (l2 & 3);
                    act = row < nrows;
                }
                double t = -INFINITY;
                int ta = 0x7fffffff;
                if (act) {
                    const int2 me = sRowLC[row];
                    const RowConst<AI> rc = make_row<AI>(me.y, me.x, p.alpha_int, p.alpha);
                    int cc[8], ll[8];
                    double pc[8], gg[8], lg[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        cc[c] = __ldg(p.C + c0 + c);
                        ll[c] = __ldg(p.L + c0 + c);
                        pc[c] = __ldcg(p.P + c0 + c);
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        gg[c] = __ldg(p.gtab + (rc.cjx - cc[c]));
                        lg[c] = __ldg(p.ltab + (rc.lj - ll[c]));
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const double sx = AI ? u32_to_double(rc.cjx - cc[c]) : __dsub_rn(rc.aj, u32_to_double(cc[c]));
                        const double v = __dadd_rn(__dsub_rn(gg[c], __dmul_rn(sx, lg[c])), pc[c]);
                        if (v > t) { t = v; ta = c0 + c; }
                    }
                }
                // rectangles of one pass may share rows: lanes with the same row update it one after the other
                const unsigned same = __match_any_sync(0xffffffffu, act ? row : -1 - lane);
                const int turn = __popc(same & ((1u << lane) - 1u));
                const int turns = __reduce_max_sync(0xffffffffu, turn);
                for (int r = 0; r <= turns; ++r) {
                    if (act && turn == r) {
                        const double cur = sWV[warp * XP_RB + row];
                        if (t > cur || (t == cur && ta < sWA[warp * XP_RB + row])) { sWV[warp * XP_RB + row] = t; sWA[warp * XP_RB + row] = ta; }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        tl2 += xp_clock<PROF>() - tq;
        if (tid == 0) { sCnt[0] = 0; sCnt[1] = 0; }
        __syncthreads();
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
        __stcg(p.farV + ((size_t)b * XP_GT + g) * XP_RB + tid, best);
        __stcg(p.farA + ((size_t)b * XP_GT + g) * XP_RB + tid, arg);
    }
    if (tid == 0 && evaluated) atomicAdd(p.far_cells, evaluated);
    xp_far_finish(p, b);
    if (tid == 0) {
        const u64 work = (u64)(xp_clock<PROF>() - tf1);
        atomicAdd(p.prof + XQ_F_WAIT, (u64)(tf1 - tf0));
        atomicAdd(p.prof + XQ_F_WORK, work);
        atomicAdd(p.prof + XQ_F_COUNT, 1ull);
        atomicMax(p.prof + XQ_F_MAX, work);
        atomicAdd(p.prof + XQ_F_L0, (u64)tl0);
        atomicAdd(p.prof + XQ_F_L1, (u64)tl1);
        atomicAdd(p.prof + XQ_F_L23, (u64)tl2);
        atomicAdd(p.prof + XQ_F_HEAD, (u64)thead);
    }
}

// R(b): the records of the finished row block b -- its four 32-column records, its 128-column record, its anchors, its
// largest |P| -- and then done_block = b + 1, in block order.  Everything here needs only final P / prev in global memory
// (p_block), so it does not have to run on the diagonal's SM: there the book-keeping warp spent 2.7 k cycles per step on the
// 32-column record alone and was the last at the step barrier in three steps out of four
// (profiles/r02_exact_dp_v9_notes.txt).
__device__ void xp_r_task(const XpParams &p, int b)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    xp_wait_cta(p.p_block, b + 1);                                      // P and prev of blocks 0 .. b are final
    const double tilt_c = __ldg(p.consts + 1), tilt_l = __ldg(p.consts + 2);
    const int c0 = 1 + XP_RB * b;                                       // (a block with an R task is never the last: 128 columns)
    if (warp < 4) {
        const int col = c0 + 32 * warp + lane;
        const int cme = __ldg(p.C + col), lme = __ldg(p.L + col);
        const double P = __ldcg(p.P + col);
        XpRec32 *rec = p.rec32 + (4 * b + warp);
        fit_column_record(cme, lme, P, &rec->r, tilt_c, tilt_l);
        if ((lane & 7) == 0) { rec->sub[lane >> 3][0] = cme; rec->sub[lane >> 3][2] = lme; }
        if ((lane & 7) == 7) { rec->sub[lane >> 3][1] = cme; rec->sub[lane >> 3][3] = lme; }
        double pm = fabs(P);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) pm = fmax(pm, __shfl_xor_sync(0xffffffffu, pm, off));
        // (non-negative doubles order like their bit patterns)
        if (lane == 0) atomicMax(reinterpret_cast<u64 *>(p.pmax), (u64)__double_as_longlong(pm));
    } else if (warp == 4) {
        int c4[4], l4[4];
        double p4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            c4[q] = __ldg(p.C + c0 + lane + 32 * q);
            l4[q] = __ldg(p.L + c0 + lane + 32 * q);
            p4[q] = __ldcg(p.P + c0 + lane + 32 * q);
        }
        fit_column_record128(c4, l4, p4, p.rec128 + b, tilt_c, tilt_l);
    } else if (warp == 5) {
        // anchors: the block's last row e and the last seven split points of the best segmentation ending there (e -> prev[e]
        // -> prev[prev[e]] ...): the rows ahead most likely continue one of these, so "best up to the anchor, then one
        // segment" bounds their maxima from below tightly
        const int e = XP_RB * (b + 1);
        int a = e, mine = e;
        for (int t = 1; t < 8; ++t) {                                   // (every lane walks the same chain: uniform loads)
            if (a > 0) a = __ldcg(p.prev + a);
            if (lane == t) mine = a;
        }
        if (lane < 8) {
            mine = min(max(mine, 0), e);
            XpAnchors *an = p.anchors + b;
            an->idx[lane] = mine;
            an->L[lane] = __ldg(p.L + mine);
            an->C[lane] = __ldg(p.C + mine);
            an->P[lane] = mine > 0 ? __ldcg(p.P + mine) : 0.0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        while (xp_ld_flag(p.done_block) < b) __nanosleep(32);           // (R(b - 1) is earlier in the task list)
        __threadfence();
        *reinterpret_cast<volatile int *>(p.done_block) = b + 1;
    }
}

// N(b, g): rows of block b against 16 columns of the first block of the band, [F_b + 16 g, F_b + 16 g + 16) -- the block
// that has just been published when block b - lag + 2 starts.  Too close to the diagonal for the bounds of the F tasks to
// be worth their latency, too far for the chain to need it soon: its 128 x 128 cells are simply evaluated, in the
// reference's operation order, by eight worker CTAs, and reach the diagonal with the far results.  This takes a third of
// the columns (at lag 3) out of the sweep that the diagonal's helper warps stream through one SM.
template <bool AI>
__device__ void xp_n_task(const XpParams &p, int b, int g, unsigned char *smem)
{
    double *sV = reinterpret_cast<double *>(smem);                      // [2][128]
    int *sA = reinterpret_cast<int *>(sV + 2 * XP_RB);                  // [2][128]
    const int tid = threadIdx.x;
    const int N = p.N;
    const int r0 = 1 + XP_RB * b, nrows = min(XP_RB, N - r0);

    const int r = tid & (XP_RB - 1), h = tid >> 7;                      // row; which 8 of the slice's 16 columns
    double best = -INFINITY;
    int arg = 0x7fffffff;
    // Column blocks b - lag + 1 .. b - lag + nb (with lag 5 and three of them the far tasks get four blocks of time, the band of
    // the diagonal stays two blocks), ascending columns within a thread.  Block by block: the task has been waiting since long
    // before, so all but its last block are evaluated before that one is released -- and everything that does not depend on
    // P (the columns' counts, the table gathers) is fetched BEFORE the wait: behind it only P is loaded (one round trip to L2,
    // not three: this task's last block and the merge decide when the diagonal gets the block's far results)
    const int j = min(r0 + r, N - 1);
    const RowConst<AI> row = make_row<AI>(__ldg(p.C + j), __ldg(p.L + j), p.alpha_int, p.alpha);
    for (int nq = 0; nq < p.nb; ++nq) {
        const int c0 = 1 + XP_RB * (b - p.lag + 1 + nq) + 16 * g + 8 * h;   // (negative in the first blocks: column 0 is "the last of block -1")
        const bool mine = r < nrows && c0 + 7 >= 0;
        double gq[8], lq[8];
        int ci[8];
        if (mine) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {                               // all gathers in flight
                const int col = max(c0 + u, 0);
                ci[u] = __ldg(p.C + col);
                gq[u] = __ldg(p.gtab + (row.cjx - ci[u]));
                lq[u] = __ldg(p.ltab + (row.lj - __ldg(p.L + col)));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) gq[u] = __dsub_rn(gq[u], __dmul_rn(AI ? u32_to_double(row.cjx - ci[u]) : __dsub_rn(row.aj, u32_to_double(ci[u])), lq[u]));
        }
        xp_wait_cta(p.p_block, b - p.lag + 2 + nq);                     // P of column block b - lag + 1 + nq is final (no records needed)
        if (!mine) continue;
        double pc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) pc[u] = __ldcg(p.P + max(c0 + u, 0));
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double t = __dadd_rn(gq[u], pc[u]);
            if (c0 + u >= 0 && t > best) { best = t; arg = c0 + u; }
        }
    }
    sV[h * XP_RB + r] = best;
    sA[h * XP_RB + r] = arg;
    __syncthreads();
    if (tid < nrows) {
        double v = sV[tid];
        int a = sA[tid];
        const double v1 = sV[XP_RB + tid];                             // first maximum = the smaller column on a tie
        const int a1 = sA[XP_RB + tid];
        if (v1 > v || (v1 == v && a1 < a)) { v = v1; a = a1; }
        __stcg(p.farV + ((size_t)b * XP_GT + XP_G + g) * XP_RB + tid, v);
        __stcg(p.farA + ((size_t)b * XP_GT + XP_G + g) * XP_RB + tid, a);
    }
    xp_far_finish(p, b);
}

// constants of the bound, once per launch (so that no far task starts with a chain of dependent loads)
template <bool AI>
__global__ void xp_consts_kernel(const int32_t *__restrict__ L, const int32_t *__restrict__ C, int N, const double *__restrict__ gtab,
                                 const double *__restrict__ ltab, int alpha_int, double alpha, double *out)
{
    const int zC = C[N - 1], zL = L[N - 1];
    out[0] = fabs(gtab[zC + (AI ? alpha_int : 0)]) + ((double)zC + alpha) * fabs(ltab[zL]) + 1.0;
    out[1] = (double)zC + alpha;
    out[2] = (double)zL;
}

template <bool AI, bool RING, bool PROF>
__global__ void __launch_bounds__(XP_THREADS, 1)
exact_pruned_kernel(XpParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int sTask;
    if (blockIdx.x == 0) {
        xp_diagonal<AI, RING, PROF>(p, smem);
        return;
    }
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sTask = (int)atomicAdd(p.task_counter, 1u);
        __syncthreads();
        const int t = sTask;
        if (t >= p.n_tasks) return;
        const int2 task = __ldg(p.tasks + t);
        if ((task.x & 1) == 0) xp_s_task<AI, PROF>(p, task.x >> 1, task.y, smem);
        else if (task.y < XP_G) xp_f_task<AI, PROF>(p, task.x >> 1, task.y, smem);
        else if (task.y < XP_GT) xp_n_task<AI>(p, task.x >> 1, task.y - XP_G, smem);
        else xp_r_task(p, task.x >> 1);
    }
}

// S tasks run XP_SAHEAD blocks ahead of the F tasks; F(b) sits where the diagonal finishes block b - lag.
std::vector<int2> build_tasks(int nB, int lag, int nb)
{
    std::vector<int2> t;
    auto push_s = [&](int b) { if (b < nB) for (int q = 0; q < XP_SQ; ++q) t.push_back(make_int2(b << 1, q)); };
    // F slices first (they are released one block earlier), then the N slices of the same row block (sub index >= XP_G)
    auto push_f = [&](int b) {
        if (b >= nB) return;
        if (b >= lag - 1) for (int g = 0; g < XP_G; ++g) t.push_back(make_int2((b << 1) | 1, g));
        if (xp_has_n(b, lag, nb)) for (int g = XP_G; g < XP_GT; ++g) t.push_back(make_int2((b << 1) | 1, g));
    };
    for (int b = 0; b < XP_SAHEAD; ++b) push_s(b);
    for (int b = 0; b < lag - 1; ++b) push_f(b);          // (two N blocks: the row blocks that have N tasks before they have F tasks)
    for (int tt = 0; tt < nB; ++tt) {
        if (tt >= 1 && tt + lag - 1 < nB) t.push_back(make_int2(((tt - 1) << 1) | 1, XP_GT));     // R(tt - 1): F(tt + lag - 1) waits for it
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
    p.dbg = getenv("PASIO_XD_DBG") ? atoi(getenv("PASIO_XD_DBG")) : 0;
    if (getenv("PASIO_XD_PROF")) p.dbg |= 128;                 // per-step-type counters (one RED per role and step)
    p.npad = (int)((N + 127) & ~(i64)127);
    p.nb = ctx->tune[PASIO_TUNE_EXACT_NBLOCK] < 0 ? 0 : (ctx->tune[PASIO_TUNE_EXACT_NBLOCK] > lag - 2 ? lag - 2 : ctx->tune[PASIO_TUNE_EXACT_NBLOCK]);   // the band keeps >= 2 blocks
    p.DB = XP_RB * (lag - p.nb);                                 // distances kept per row block: the band is one block narrower with N tasks
    const std::vector<int2> tasks = build_tasks(p.nB, lag, p.nb);
    p.n_tasks = (int)tasks.size();

    // every row block its own slab of self scores while that fits XP_LINEAR_BYTES (2 GB: N <= 680 000 at lag 3), else a ring
    // a ring of XP_RING slabs (25 MB at lag 3) stays in L2; one slab per block (307 / 614 MB for configs 1 / 3) was measured
    // slower (profiles/r02_exact_dp_v7_*: 9.6 vs 8.9 ms): its slabs are written back to DRAM before the diagonal reads them
    const bool ring = ctx->tune[PASIO_TUNE_EXACT_RING] != 0 || (size_t)p.nB * p.DB * XP_RB * 8 > XP_LINEAR_BYTES;
    const int slots = ring && p.nB > XP_RING ? XP_RING : p.nB;
    p.s_slots = slots;
    const size_t ring_bytes = (size_t)slots * p.DB * XP_RB * 8;
    const size_t far_bytes = (size_t)(XP_GT + 1) * XP_RB * p.nB * 12;     // slices + the merged results
    const size_t rec_bytes = ((size_t)p.nSteps + 1) * sizeof(XpRec32) + ((size_t)p.nB + 1) * (sizeof(CoarseRec) + sizeof(XpAnchors));
    const size_t flag_ints = (size_t)3 * p.nB + 8 + 128 + 8;     // + 64 u64 profile counters + 4 doubles of constants
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
    p.p_block = flags + 6;
    p.task_counter = reinterpret_cast<unsigned *>(flags + 5);
    p.prof = reinterpret_cast<u64 *>(flags + 8);                // 64 x 8 bytes
    double *consts = reinterpret_cast<double *>(flags + 8 + 128);   // 4 x 8 bytes
    p.consts = consts;
    p.s_ready = flags + 8 + 128 + 8;
    p.far_ready = p.s_ready + p.nB;
    p.far_done = p.far_ready + p.nB;
    p.tasks = ctx->xpTasks.as<int2>();
    p.farV = ctx->dpPart.as<double>();
    p.farVm = p.farV + (size_t)XP_GT * XP_RB * p.nB;            // (8-byte data first: the bulk copies need 16-byte aligned rows)
    p.farA = reinterpret_cast<int *>(p.farVm + (size_t)XP_RB * p.nB);
    p.farAm = p.farA + (size_t)XP_GT * XP_RB * p.nB;
    p.rec32 = ctx->xpRec.as<XpRec32>();
    p.rec128 = reinterpret_cast<CoarseRec *>(p.rec32 + p.nSteps + 1);
    p.anchors = reinterpret_cast<XpAnchors *>(p.rec128 + p.nB + 1);
    p.gtab = ctx->tab[AI ? PASIO_TAB_LGAMMA : PASIO_TAB_LGAMMA_ALPHA].as<double>();
    p.ltab = ctx->tab[PASIO_TAB_LOG].as<double>();
    p.alpha_int = (int)ctx->alpha_int;
    p.alpha = ctx->alpha;
    p.pen = ctx->pen;

    xp_consts_kernel<AI><<<1, 1, 0, ctx->stream>>>(p.L, p.C, p.N, p.gtab, p.ltab, p.alpha_int, p.alpha, consts);
    const size_t smem = xp_diag_smem() > xp_worker_smem() ? xp_diag_smem() : xp_worker_smem();
    static const bool with_counters = getenv("PASIO_XD_PROF") != nullptr || getenv("PASIO_XD_DBG") != nullptr;
    void (*kern)(XpParams) = with_counters ? (ring ? exact_pruned_kernel<AI, true, true> : exact_pruned_kernel<AI, false, true>)
                                           : (ring ? exact_pruned_kernel<AI, true, false> : exact_pruned_kernel<AI, false, false>);
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, XP_THREADS, smem));
    if (per_sm < 1) {
        cudaFuncAttributes fa;
        memset(&fa, 0, sizeof fa);
        cudaFuncGetAttributes(&fa, kern);
        return pasio_fail(ctx, PASIO_E_CUDA, "pruned exact DP kernel cannot be resident (smem %zu, regs %d, static smem %zu, max dynamic %d, "
                          "ring %d, N %d)", smem, fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes, (int)ring, p.N);
    }
    int workers = ctx->sm_count - 1;                            // one CTA per SM: the diagonal has an SM to itself
    if (workers > p.n_tasks) workers = p.n_tasks;
    if (workers < 1) workers = 1;
    void *args[] = {&p};
    {
        TimingScope ts(ctx, TF_EXACT_DP);
        // cooperative launch = all CTAs co-resident, which the flag waits rely on
        CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void *)kern, dim3(1 + workers), dim3(XP_THREADS), args, smem, ctx->stream));
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
    static const bool prof = getenv("PASIO_XD_PROF") != nullptr;
    if (prof) {
        u64 h[64];
        cudaMemcpy(h, p.prof, sizeof h, cudaMemcpyDeviceToHost);
        const double st = (double)p.nSteps;
        fprintf(stderr, "[xp_prof] N=%d steps=%d lag=%d workers=%d | cycles/step: diag %.0f = chain %.0f + barrier wait %.0f | book-keeping warp: rec %.0f (gaps %.0f %.0f) "
                        "s-wait %.0f far %.0f bar %.0f | sweeping warp: mid %.0f tile %.0f s-wait %.0f far %.0f bar %.0f\n",
                p.N, p.nSteps, lag, workers, h[XQ_DIAG_TOTAL] / st, h[XQ_CHAIN] / st, h[XQ_CHAIN_BAR] / st, h[XQ_H6_REC] / st, h[XQ_H6_MID] / st,
                h[XQ_H6_TILE] / st, h[XQ_H6_SWAIT] / st, h[XQ_H6_FAR] / st, h[XQ_H6_BAR] / st, h[XQ_H0_MID] / st, h[XQ_H0_TILE] / st,
                h[XQ_H0_SWAIT] / st, h[XQ_H0_FAR] / st, h[XQ_H0_BAR] / st);
        fprintf(stderr, "[xp_prof] mid cycles/step of the six sweeping warps (distance chunks 0 .. 5): %.0f %.0f %.0f %.0f %.0f %.0f\n",
                h[14] / st, h[15] / st, h[27] / st, h[28] / st, h[29] / st, h[30] / st);
        fprintf(stderr, "[xp_prof] wait at the step barrier (shows between the clock read behind it and the next step's: a clock read does not wait for BAR): chain warp %.0f, sweeping warp %.0f, book-keeping warp %.0f cycles/step\n",
                h[31] / st, h[XQ_H6_MID] / st, h[XQ_H6_TILE] / st);
        const double sq = st / 4.0;
        fprintf(stderr, "[xp_prof] cycles from the top of a step to the step barrier, by step type k & 3 = 0 1 2 3: chain warp %.0f %.0f %.0f %.0f | sweeping warp %.0f %.0f %.0f %.0f | "
                        "book-keeping warp %.0f %.0f %.0f %.0f | whole step %.0f %.0f %.0f %.0f | blocking looks at the self-score flags %llu\n",
                h[32] / sq, h[33] / sq, h[34] / sq, h[35] / sq, h[36] / sq, h[37] / sq, h[38] / sq, h[39] / sq, h[40] / sq, h[41] / sq, h[42] / sq, h[43] / sq,
                h[48] / sq, h[49] / sq, h[50] / sq, h[51] / sq, (unsigned long long)h[44]);
        const double ns = (double)(h[XQ_S_COUNT] ? h[XQ_S_COUNT] : 1), nf = (double)(h[XQ_F_COUNT] ? h[XQ_F_COUNT] : 1);
        fprintf(stderr, "[xp_prof] S tasks %llu: work %.0f cyc, wait %.0f | F tasks %llu: wait %.0f, work %.0f (max %llu) = head %.0f + L0 %.0f + L1 %.0f + L2/3 %.0f\n",
                (unsigned long long)h[XQ_S_COUNT], h[XQ_S_WORK] / ns, h[XQ_S_WAIT] / ns, (unsigned long long)h[XQ_F_COUNT], h[XQ_F_WAIT] / nf,
                h[XQ_F_WORK] / nf, (unsigned long long)h[XQ_F_MAX], h[XQ_F_HEAD] / nf, h[XQ_F_L0] / nf, h[XQ_F_L1] / nf, h[XQ_F_L23] / nf);
    }
    return PASIO_OK;
}

}  // namespace

int launch_exact_dp_pruned(pasio_ctx *ctx, i64 N, int lag)
{
    PASIO_TRY(pasio_reserve(ctx, ctx->dpP, (size_t)N * 8));
    PASIO_TRY(pasio_reserve(ctx, ctx->dpPrev, (size_t)N * 4));
    return ctx->alpha_is_int ? run_exact_pruned<true>(ctx, N, lag) : run_exact_pruned<false>(ctx, N, lag);
}

extern "C" int pasio_exact_task_plan(int64_t n_candidates, int lag, int nblock, int32_t *triples, int64_t cap, int64_t *n_tasks)
{
    if (n_candidates < 2 || lag < 3 || lag > 5 || !n_tasks) return PASIO_E_ARG;
    const int nB = (int)((n_candidates - 1 + XP_RB - 1) / XP_RB);
    const int nb = nblock < 0 ? 0 : (nblock > lag - 2 ? lag - 2 : nblock);      // as run_exact_pruned clamps it
    const std::vector<int2> tasks = build_tasks(nB, lag, nb);
    *n_tasks = (int64_t)tasks.size();
    if (!triples || cap < (int64_t)tasks.size()) return PASIO_E_ARG;
    for (size_t i = 0; i < tasks.size(); ++i) {
        const int b = tasks[i].x >> 1, y = tasks[i].y;
        const int kind = (tasks[i].x & 1) == 0 ? 0 : (y < XP_G ? 1 : (y < XP_GT ? 2 : 3));
        triples[3 * i] = b;
        triples[3 * i + 1] = kind;
        triples[3 * i + 2] = kind == 2 ? y - XP_G : (kind == 3 ? 0 : y);
    }
    return PASIO_OK;
}
