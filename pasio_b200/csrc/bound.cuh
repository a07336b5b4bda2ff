// The exact branch-and-bound shared by the batched window kernel (K4, window_dp.cu) and the whole-contig exact DP
// (K3, exact_pruned.cu): records of finished column blocks, the tilted corner bound, the tilt fit.
//
// For rows j in [j0, j1] and columns i in [i0, i1] every DP cell value (dp_core.cuh)
//     t_ij = F(u_ij, len_ij) + P_i ,  F(u, len) = G[u (+alpha)] - s*Lg[len],  u_ij = C_j - C_i,  len_ij = L_j - L_i
// satisfies, for ANY real a, b (the "tilt"),
//     t_ij - LB_j = [P_i + a*C_i + b*L_i] + [F(u_ij, len_ij) + a*u_ij + b*len_ij] - [LB_j + a*C_j + b*L_j]
//                <= max_i [P_i + a*C_i + b*L_i]  +  max_box [F(u, len) + a*u + b*len]  -  min_j [LB_j + a*C_j + b*L_j]
// where the box is [C_j0 - C_i1, C_j1 - C_i0] x [L_j0 - L_i1, L_j1 - L_i0].  F + a*u + b*len is convex in u for fixed
// len (lgamma is convex, the rest is linear) and convex in len for fixed u (-s*log(len + beta) with s >= 0), so its
// maximum over the box is at one of the 4 corners.  If the right-hand side (plus a delta that covers table and
// rounding errors) is negative and LB_j is a lower bound of row j's maximum, no cell of the rectangle holds the
// arg-max or a tie: skipping it leaves P, prev and the back-trace bit-identical.
#pragma once
#include "dp_core.cuh"

// One finished block of 32 columns [1+32b, 32+32b], as the far pass sees it.
struct __align__(16) CoarseRec {
    int c_first, c_last, l_first, l_last;   // C and L of its first / last column
    double a, b;                            // tilt: P_i + a*C_i + b*L_i is nearly constant over the block
    double mpt;                             // max_i (P_i + a*C_i + b*L_i) over the block, raised by the tilt's rounding slack
    double mpt8[4];                         // the same over each 8-column sub-block
    double pad;
};
static_assert(sizeof(CoarseRec) == 80, "CoarseRec layout");

// The four table values a box bound needs, separated from the arithmetic so that a caller can have the gathers of many
// boxes in flight before it consumes any (exact_pruned.cu).
struct BoxCorners { double g_lo, g_hi, l_lo, l_hi; };

template <bool AI>
__device__ __forceinline__ BoxCorners box_corner_loads(int u_lo, int u_hi, int len_lo, int len_hi,
                                                       const double *__restrict__ gtab, const double *__restrict__ ltab, int alpha_int)
{
    BoxCorners c;
    c.g_lo = __ldg(gtab + (AI ? u_lo + alpha_int : u_lo));
    c.g_hi = __ldg(gtab + (AI ? u_hi + alpha_int : u_hi));
    c.l_lo = __ldg(ltab + len_lo);
    c.l_hi = __ldg(ltab + len_hi);
    return c;
}

__device__ __forceinline__ double tilted_box_max_of(const BoxCorners &c, int u_lo, int u_hi, int len_lo, int len_hi, double a, double b,
                                                    double alpha)
{
    const double ud_lo = u32_to_double(u_lo), ud_hi = u32_to_double(u_hi);
    const double s_lo = ud_lo + alpha, s_hi = ud_hi + alpha;
    const double ta_lo = a * ud_lo, ta_hi = a * ud_hi;
    const double tb_lo = b * u32_to_double(len_lo), tb_hi = b * u32_to_double(len_hi);
    const double f00 = (c.g_lo - s_lo * c.l_lo) + (ta_lo + tb_lo);
    const double f01 = (c.g_lo - s_lo * c.l_hi) + (ta_lo + tb_hi);
    const double f10 = (c.g_hi - s_hi * c.l_lo) + (ta_hi + tb_lo);
    const double f11 = (c.g_hi - s_hi * c.l_hi) + (ta_hi + tb_hi);
    return fmax(fmax(f00, f01), fmax(f10, f11));
}

template <bool AI>
__device__ __forceinline__ double tilted_box_max(int u_lo, int u_hi, int len_lo, int len_hi, double a, double b,
                                                 const double *__restrict__ gtab, const double *__restrict__ ltab,
                                                 int alpha_int, double alpha)
{
    const BoxCorners c = box_corner_loads<AI>(u_lo, u_hi, len_lo, len_hi, gtab, ltab, alpha_int);
    return tilted_box_max_of(c, u_lo, u_hi, len_lo, len_hi, a, b, alpha);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}


// Tilt of a block of finished columns: least squares of -P on (C, L).  The fit only steers how tight the bound
// is (any finite a, b is valid), so its sums run in float.  sums: centred second moments over the block.
__device__ __forceinline__ void solve_tilt(float cxx, float cyy, float cxy, float cxp, float cyp, double &a, double &b)
{
    a = 0.0;
    b = 0.0;
#ifndef PASIO_NO_LS
    const float det = cxx * cyy - cxy * cxy;
    float fa = 0.f, fb = 0.f;
    if (det > 1e-4f * cxx * cyy) {
        fa = -(cxp * cyy - cyp * cxy) / det;
        fb = -(cyp * cxx - cxp * cxy) / det;
    } else if (cxx > 0.f) {
        fa = -cxp / cxx;
    } else if (cyy > 0.f) {
        fb = -cyp / cyy;
    }
    if (fabsf(fa) < 1e30f && fabsf(fb) < 1e30f) { a = (double)fa; b = (double)fb; }   // NaN / inf: no tilt
#endif
}

// Record of a finished block of 32 columns, one warp, lane = column: tilt, tilted maxima, end points.
__device__ __forceinline__ void fit_column_record(int C, int L, double P, CoarseRec *rec, double tilt_scale_c,
                                                  double tilt_scale_l)
{
    const int lane = threadIdx.x & 31;
    const int c_first = __shfl_sync(0xffffffffu, C, 0), c_last = __shfl_sync(0xffffffffu, C, 31);
    const int l_first = __shfl_sync(0xffffffffu, L, 0), l_last = __shfl_sync(0xffffffffu, L, 31);
    const double p_first = __shfl_sync(0xffffffffu, P, 0);
    double a, b;                                   // fit  -P ~ a*C + b*L + const
    {
        const float x = (float)(C - c_first), y = (float)(L - l_first), p = (float)(P - p_first);
        const float inv_n = 1.0f / 32.0f;
        const float sx = warp_sum(x), sy = warp_sum(y), sp = warp_sum(p);
        const float xc = x - sx * inv_n, yc = y - sy * inv_n, pc = p - sp * inv_n;
        const float cxx = warp_sum(xc * xc), cyy = warp_sum(yc * yc), cxy = warp_sum(xc * yc);
        const float cxp = warp_sum(xc * pc), cyp = warp_sum(yc * pc);
        solve_tilt(cxx, cyy, cxy, cxp, cyp, a, b);
    }
    double m = P + (a * u32_to_double(C) + b * u32_to_double(L));
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    double m32 = m;
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) m32 = fmax(m32, __shfl_xor_sync(0xffffffffu, m32, off));
    const double slack = (fabs(a) * tilt_scale_c + fabs(b) * tilt_scale_l) * 5.684341886080802e-14;   // 2^-44
    if ((lane & 7) == 0) rec->mpt8[lane >> 3] = m + slack;
    if (lane == 0) {
        *reinterpret_cast<int4 *>(rec) = make_int4(c_first, c_last, l_first, l_last);
        rec->a = a;
        rec->b = b;
        rec->mpt = m32 + slack;
    }
}

// The same for a block of 128 columns (exact_pruned.cu, coarsest level), one warp: lane holds columns lane, lane+32,
// lane+64, lane+96 of the block.  mpt8[q] = tilted maximum of the q-th 32-column group under the block's tilt.
__device__ __forceinline__ void fit_column_record128(const int (&C)[4], const int (&L)[4], const double (&P)[4], CoarseRec *rec,
                                                     double tilt_scale_c, double tilt_scale_l)
{
    const int lane = threadIdx.x & 31;
    const int c_first = __shfl_sync(0xffffffffu, C[0], 0), c_last = __shfl_sync(0xffffffffu, C[3], 31);
    const int l_first = __shfl_sync(0xffffffffu, L[0], 0), l_last = __shfl_sync(0xffffffffu, L[3], 31);
    const double p_first = __shfl_sync(0xffffffffu, P[0], 0);
    double a, b;
    {
        float x[4], y[4], p[4], sx = 0.f, sy = 0.f, sp = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x[k] = (float)(C[k] - c_first);
            y[k] = (float)(L[k] - l_first);
            p[k] = (float)(P[k] - p_first);
            sx += x[k];
            sy += y[k];
            sp += p[k];
        }
        const float inv_n = 1.0f / 128.0f;
        sx = warp_sum(sx) * inv_n;
        sy = warp_sum(sy) * inv_n;
        sp = warp_sum(sp) * inv_n;
        float cxx = 0.f, cyy = 0.f, cxy = 0.f, cxp = 0.f, cyp = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xc = x[k] - sx, yc = y[k] - sy, pc = p[k] - sp;
            cxx += xc * xc;
            cyy += yc * yc;
            cxy += xc * yc;
            cxp += xc * pc;
            cyp += yc * pc;
        }
        solve_tilt(warp_sum(cxx), warp_sum(cyy), warp_sum(cxy), warp_sum(cxp), warp_sum(cyp), a, b);
    }
    const double slack = (fabs(a) * tilt_scale_c + fabs(b) * tilt_scale_l) * 5.684341886080802e-14;   // 2^-44
    double mall = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double m = P[k] + (a * u32_to_double(C[k]) + b * u32_to_double(L[k]));
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (lane == 0) rec->mpt8[k] = m + slack;
        mall = fmax(mall, m);
    }
    if (lane == 0) {
        *reinterpret_cast<int4 *>(rec) = make_int4(c_first, c_last, l_first, l_last);
        rec->a = a;
        rec->b = b;
        rec->mpt = mall + slack;
    }
}
