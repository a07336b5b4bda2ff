// The DP cell and the in-CTA block step shared by the batched window kernel (K4) and the
// diagonal step of the exact kernel (K3).
//
// One cell restates, in the reference's operation order,
//   LogMarginalLikelyhood{Int,Real}AlphaComputer.all_suffixes_self_score
//     /root/reference/src/pasio/log_marginal_likelyhood.py:105-115 (int alpha), :121-132 (real alpha)
//   + the row update of SquareSplitter.split_without_normalizations
//     /root/reference/src/pasio/splitters/square_splitter.py:84-94
//       s    = (alpha + C_j) - C_i
//       self = G[s] - s * Lg[L_j - L_i]          (un-fused DMUL then DADD, like numpy's two ufunc loops)
//       t    = self + P_i
//       P_j  = max_i t  + segment_creation_cost,  prev_j = first index of the maximum (np.argmax)
// G and Lg are host-built tables (numpy/scipy values), so the device does no transcendental math.
#pragma once
#include "common.cuh"

// Per-row constants.  AI = integer alpha (table index includes alpha, s is that integer);
// real alpha: table index is the raw count, s = (alpha + C_j) - C_i in float64 with the
// reference's two roundings.
template <bool AI>
struct RowConst {
    int cjx;      // AI: C_j + alpha ; else C_j
    int lj;       // L_j
    double aj;    // !AI: fl(alpha + C_j)
};

template <bool AI>
__device__ __forceinline__ RowConst<AI> make_row(int cj, int lj, int alpha_int, double alpha)
{
    RowConst<AI> r;
    r.lj = lj;
    if (AI) {
        r.cjx = cj + alpha_int;
        r.aj = 0.0;
    } else {
        r.cjx = cj;
        r.aj = __dadd_rn(alpha, u32_to_double(cj));
    }
    return r;
}

// self score of segment [i, j): G[s] - s * Lg[len]
template <bool AI>
__device__ __forceinline__ double self_score(int ci, int li, const RowConst<AI> &r,
                                             const double *__restrict__ gtab, const double *__restrict__ ltab)
{
    const int idx = r.cjx - ci;
    const int len = r.lj - li;
    const double g = __ldg(gtab + idx);
    const double lg = __ldg(ltab + len);
    const double s = AI ? u32_to_double(idx) : __dsub_rn(r.aj, u32_to_double(ci));
    return __dsub_rn(g, __dmul_rn(s, lg));
}

// Sweep columns i0+phase, i0+phase+stride, ... < i1 for this lane's row; strict '>' keeps the first
// maximum among the lane's own (ascending) columns.
// sLC[i] = (L_i, C_i), sP[i] = P_i (shared memory; lanes of one column phase read one address).
// The table gathers are the long-latency part: each batch of U cells first issues all 2*U
// gathers, then does the arithmetic, so a warp keeps 2*U loads in flight.
template <bool AI, int U>
__device__ __forceinline__ void sweep_columns(int i0, int i1, int phase, int stride, const int2 *sLC,
                                              const double *sP, const RowConst<AI> &r,
                                              const double *__restrict__ gtab, const double *__restrict__ ltab,
                                              double &best, int &arg)
{
    int i = i0 + phase;
    for (; i + (U - 1) * stride < i1; i += U * stride) {
        double g[U], lg[U];
        int sx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int2 a = sLC[i + u * stride];
            const int idx = r.cjx - a.y;
            g[u] = __ldg(gtab + idx);
            lg[u] = __ldg(ltab + (r.lj - a.x));
            sx[u] = AI ? idx : a.y;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double s = AI ? u32_to_double(sx[u]) : __dsub_rn(r.aj, u32_to_double(sx[u]));
            const double t = __dadd_rn(__dsub_rn(g[u], __dmul_rn(s, lg[u])), sP[i + u * stride]);
            if (t > best) { best = t; arg = i + u * stride; }
        }
    }
    for (; i < i1; i += stride) {
        const int2 a = sLC[i];
        const double t = __dadd_rn(self_score<AI>(a.y, a.x, r, gtab, ltab), sP[i]);
        if (t > best) { best = t; arg = i; }
    }
}

// Combine the (max, arg) of the CPW lanes that swept interleaved column phases of one row:
// larger value wins, equal values keep the smaller column index (np.argmax's first maximum).
template <int RPW>
__device__ __forceinline__ void merge_column_phases(double &best, int &arg)
{
#pragma unroll
    for (int off = RPW; off < 32; off <<= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, off);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
}

constexpr int DP_JB = 32;   // rows resolved per block step (one per lane)

// One block step over rows [jb, jb+32) of a candidate list held in shared memory.
//   rectangle: columns [col0, jb) split across the CTA's warps (lane = row);
//   triangle : columns [jb, j) -- the 32x32 self scores are computed by all warps, then warp 0
//              resolves the 32 rows in order, broadcasting each finished P by shuffle.
// init_best/init_arg (warp 0 only, per lane) seed the running maximum with what earlier columns
// (outside [col0, jb)) contributed; pass -inf / 0 when there are none.
// Requires col0 <= jb, blockDim.x == NW*32.  Ends with a __syncthreads().
template <bool AI, int NW, int U, int RPW>
__device__ __forceinline__ void dp_block_step(int jb, int N, int col0, const int2 *sLC, double *sP,
                                              unsigned short *sPrev16, int *sPrev32,
                                              double *sPartV, int *sPartA, double *sTri,
                                              const double *__restrict__ gtab, const double *__restrict__ ltab,
                                              int alpha_int, double alpha, double pen,
                                              double init_best, int init_arg, int arg_offset)
{
    // A warp covers RPW rows x CPW interleaved column phases per step: the 32 gather addresses of one
    // load then span RPW candidates of rows plus CPW candidates of columns instead of 32 candidates,
    // which is what sets the number of L1 lines a gather touches (profiles/r01_window_dp_v2_*).
    constexpr int CPW = 32 / RPW;         // column phases per warp
    constexpr int NG = DP_JB / RPW;       // row groups per 32-row block
    constexpr int NQ = NW / NG;           // column chunks (warps per row group)
    static_assert(NW % NG == 0 && NQ >= 1, "warps must tile the row groups");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = warp % NG, q = warp / NG;
    const int rr = lane % RPW, cc = lane / RPW;

    // rectangle: columns [col0, jb) in NQ chunks
    {
        const int j = min(jb + grp * RPW + rr, N - 1);
        const int2 lc = sLC[j];
        const RowConst<AI> r = make_row<AI>(lc.y, lc.x, alpha_int, alpha);
        const int ncol = jb - col0;
        const int chunk = (ncol + NQ - 1) / NQ;
        const int i0 = col0 + q * chunk;
        const int i1 = min(i0 + chunk, jb);
        double best = -INFINITY;
        int arg = i0 + cc;
        sweep_columns<AI, U>(i0, i1, cc, CPW, sLC, sP, r, gtab, ltab, best, arg);
        merge_column_phases<RPW>(best, arg);
        if (cc == 0) {
            sPartV[q * 32 + grp * RPW + rr] = best;
            sPartA[q * 32 + grp * RPW + rr] = arg;
        }
    }

    // triangle self scores: pair (k, l), column jb+k, row jb+l, k < l  (lane = row here)
    {
        const int2 lc = sLC[min(jb + lane, N - 1)];
        const RowConst<AI> r = make_row<AI>(lc.y, lc.x, alpha_int, alpha);
        for (int k = warp; k < DP_JB; k += NW) {
            if (k < lane && jb + lane < N) {
                const int2 a = sLC[jb + k];
                sTri[k * DP_JB + lane] = self_score<AI>(a.y, a.x, r, gtab, ltab);
            }
        }
    }
    __syncthreads();

    if (warp == 0) {
        double best = init_best;
        int arg = init_arg;
#pragma unroll
        for (int w = 0; w < NQ; ++w) {
            double v = sPartV[w * 32 + lane];
            if (v > best) { best = v; arg = sPartA[w * 32 + lane] + arg_offset; }
        }
        double mine = 0.0;
        const int rows = min(DP_JB, N - jb);
        for (int k = 0; k < rows; ++k) {
            const double pf = __dadd_rn(best, pen);          // prefix_scores[j] = max + segment_creation_cost
            const double pk = __shfl_sync(0xffffffffu, pf, k);
            if (lane == k) mine = pf;
            if (lane > k) {
                const double t = __dadd_rn(sTri[k * DP_JB + lane], pk);
                if (t > best) { best = t; arg = jb + k + arg_offset; }
            }
        }
        if (jb + lane < N) {
            sP[jb + lane] = mine;
            if (sPrev16) sPrev16[jb + lane] = (unsigned short)arg;
            else sPrev32[jb + lane] = arg;
        }
    }
    __syncthreads();
}
