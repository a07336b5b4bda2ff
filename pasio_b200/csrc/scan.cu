// K1: per-contig coverage prefix sums as a single-pass decoupled-look-back scan, fused with
// the change-point bitmap of NotConstantReducer and the counts >= 0 validation.
//
// Replaces np.cumsum(counts) in LogMarginalLikelyhoodComputer.__init__
// (/root/reference/src/pasio/log_marginal_likelyhood.py:57), which the reference re-runs for
// every window, and the counts[:-1] != counts[1:] test of NotConstantReducer
// (splitters/constants_reducer.py:16-17).  HBM-bound: 8 B read + 8 B written per nt.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_SLABS = 8;                    // 64-element slabs per warp
constexpr int SCAN_WARP_ELEMS = 64 * SCAN_SLABS;
constexpr int SCAN_TILE = (SCAN_THREADS / 32) * SCAN_WARP_ELEMS;   // elements per CTA

// tile status word: top 2 bits = flag (0 invalid, 1 aggregate, 2 inclusive prefix), low 62 = value
__device__ __forceinline__ u64 pack_state(u64 flag, i64 v) { return (flag << 62) | (u64)v; }
__device__ __forceinline__ u64 ld_state(const u64 *p) { return *reinterpret_cast<const volatile u64 *>(p); }
__device__ __forceinline__ void st_state(u64 *p, u64 v) { *reinterpret_cast<volatile u64 *>(p) = v; }

__device__ __forceinline__ u64 spread_bits(unsigned x)   // bit k of x -> bit 2k
{
    u64 v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// Each warp owns SCAN_WARP_ELEMS consecutive elements: SCAN_SLABS slabs of 32 lanes x one 16-byte pair, so every load
// and store instruction of a warp is one fully coalesced 512-byte access.  cg[p] is the EXCLUSIVE
// prefix at p, which makes the (cg[2m], cg[2m+1]) pair an aligned 16-byte store.  The change-point
// bits of a slab come from two ballots interleaved into one 64-bit word.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_counts_kernel(const i64 *__restrict__ counts, i64 n, i64 *__restrict__ cg,
                   u64 *__restrict__ cpwords, u64 *tile_state, unsigned *tile_counter, i64 first_tile, i64 n_tiles,
                   i64 *scalars /* [0]=total, [1]=negative seen, [2]=largest count */)
{
    __shared__ unsigned s_tile;
    __shared__ i64 s_warp[SCAN_THREADS / 32];
    __shared__ i64 s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Persistent CTAs (one resident wave) take tile numbers from a per-launch counter: tiles start in issue order, so the
    // look-back never waits on an unscheduled tile, and a chr1-sized contig is 60 000 tiles but only a few hundred CTAs
    // (the CTA launch rate, not HBM, bounded the one-tile-per-CTA version: profiles/r02_scan_*).
    while (true) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    if ((i64)s_tile >= n_tiles) break;
    const i64 tile = first_tile + s_tile;
    const i64 wbase = tile * SCAN_TILE + (i64)warp * SCAN_WARP_ELEMS;     // first element of this warp's chunk

    i64 v0[SCAN_SLABS], v1[SCAN_SLABS];
    bool neg = false;
    i64 vmax = 0;
#pragma unroll
    for (int k = 0; k < SCAN_SLABS; ++k) {
        const i64 e = wbase + 64 * k + 2 * lane;
        if (e + 1 < n) {
            const longlong2 t = __ldg(reinterpret_cast<const longlong2 *>(counts + e));
            v0[k] = t.x;
            v1[k] = t.y;
        } else {
            v0[k] = (e < n) ? __ldg(counts + e) : 0;
            v1[k] = 0;
        }
        neg |= (v0[k] < 0) | (v1[k] < 0);
        vmax = max(vmax, max(v0[k], v1[k]));
    }
    if (neg) scalars[1] = 1;
    // largest count (the log-factorial table must reach it): one atomic per warp, and only when it raises the maximum
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0 && vmax > *reinterpret_cast<volatile i64 *>(scalars + 2)) atomicMax(reinterpret_cast<long long *>(scalars + 2), vmax);

    // change-point words: position p flagged when counts[p-1] != counts[p], 1 <= p <= n-1
    {
        i64 carry = (wbase > 0 && wbase - 1 < n) ? __ldg(counts + wbase - 1) : 0;   // element before the chunk
        u64 word[SCAN_SLABS];
#pragma unroll
        for (int k = 0; k < SCAN_SLABS; ++k) {
            const i64 e = wbase + 64 * k + 2 * lane;
            i64 before = __shfl_up_sync(0xffffffffu, v1[k], 1);
            if (lane == 0) before = carry;
            const bool f0 = (e >= 1) && (e <= n - 1) && (v0[k] != before);
            const bool f1 = (e + 1 <= n - 1) && (v1[k] != v0[k]);
            const unsigned b0 = __ballot_sync(0xffffffffu, f0);
            const unsigned b1 = __ballot_sync(0xffffffffu, f1);
            word[k] = spread_bits(b0) | (spread_bits(b1) << 1);
            carry = __shfl_sync(0xffffffffu, v1[k], 31);
        }
        u64 mine = 0;
#pragma unroll
        for (int k = 0; k < SCAN_SLABS; ++k)
            if (lane == k) mine = word[k];
        if (lane < SCAN_SLABS && wbase + 64 * lane <= n) cpwords[(wbase >> 6) + lane] = mine;
    }

    // warp-level inclusive scans of the pair sums, slab by slab
    i64 ex[SCAN_SLABS];
    i64 run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_SLABS; ++k) {
        const i64 pair = v0[k] + v1[k];
        i64 incl = pair;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const i64 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        ex[k] = run + incl - pair;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_warp[warp] = run;
    __syncthreads();
    i64 warp_off = 0, aggregate = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const i64 t = s_warp[w];
        if (w < warp) warp_off += t;
        aggregate += t;
    }

    // decoupled look-back over predecessor tiles (warp 0)
    if (warp == 0) {
        i64 running = 0;
        if (tile > 0) {
            if (lane == 0) st_state(tile_state + tile, pack_state(1, aggregate));
            i64 look = tile - 1;
            while (true) {
                const i64 idx = look - lane;
                u64 st;
                do {
                    st = (idx >= 0) ? ld_state(tile_state + idx) : pack_state(2, 0);
                } while (__any_sync(0xffffffffu, (st >> 62) == 0));
                const unsigned has_prefix = __ballot_sync(0xffffffffu, (st >> 62) == 2);
                const int stop = has_prefix ? (__ffs(has_prefix) - 1) : 31;
                i64 val = (lane <= stop) ? (i64)(st & 0x3fffffffffffffffull) : 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                running += val;
                if (has_prefix) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            st_state(tile_state + tile, pack_state(2, running + aggregate));
            s_prefix = running;
        }
    }
    __syncthreads();
    const i64 off = s_prefix + warp_off;
#pragma unroll
    for (int k = 0; k < SCAN_SLABS; ++k) {
        const i64 e = wbase + 64 * k + 2 * lane;
        const i64 c0 = off + ex[k];
        const i64 c1 = c0 + v0[k];
        if (e + 1 <= n) {
            *reinterpret_cast<longlong2 *>(cg + e) = make_longlong2(c0, c1);
            if (e + 1 == n) scalars[0] = c1;
        } else if (e <= n) {
            cg[e] = c0;
            if (e == n) scalars[0] = c0;
        }
    }
    }   // next tile
}

// Expand run-length intervals (bedgraph form) to the dense int64 profile.
__global__ void expand_rle_kernel(const i64 *__restrict__ starts, const i64 *__restrict__ values,
                                  i64 n_runs, i64 n, i64 *__restrict__ counts)
{
    const i64 base = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (base >= n) return;
    // last run with starts[r] <= base
    i64 lo = 0, hi = n_runs - 1;
    while (lo < hi) {
        i64 mid = (lo + hi + 1) >> 1;
        if (__ldg(starts + mid) <= base) lo = mid; else hi = mid - 1;
    }
    i64 r = lo;
    i64 next = __ldg(starts + r + 1);
    i64 val = __ldg(values + r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        i64 p = base + k;
        if (p >= n) break;
        while (p >= next) {
            ++r;
            next = __ldg(starts + r + 1);
            val = __ldg(values + r);
        }
        counts[p] = val;
    }
}

// ---- deterministic float64 prefix sums of lgamma(counts+1) (logfac_cumsum) ------------------
// Replaces np.cumsum(gammaln(counts + 1)) (log_marginal_likelyhood.py:59-60).  The reference sums
// sequentially; a parallel scan rounds differently, so this column is tolerance-only.
// Three fixed-shape passes so the result is reproducible run to run.
constexpr int FS_THREADS = 256;
constexpr int FS_ITEMS = 8;
constexpr int FS_TILE = FS_THREADS * FS_ITEMS;

__device__ __forceinline__ double block_excl_scan(double x, double *s_warp, double *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    double off = 0, tot = 0;
    for (int w = 0; w < FS_THREADS / 32; ++w) {
        double t = s_warp[w];
        if (w < warp) off += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return off + incl - x;
}

__global__ void __launch_bounds__(FS_THREADS)
logfac_tile_sums(const i64 *__restrict__ counts, i64 n, const double *__restrict__ gtab, double *tile_sum)
{
    __shared__ double s_warp[FS_THREADS / 32];
    const i64 base = (i64)blockIdx.x * FS_TILE + (i64)threadIdx.x * FS_ITEMS;
    double s = 0;
#pragma unroll
    for (int k = 0; k < FS_ITEMS; ++k)
        if (base + k < n) s += __ldg(gtab + __ldg(counts + base + k) + 1);
    double tot;
    block_excl_scan(s, s_warp, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(FS_THREADS)
logfac_scan_tile_sums(double *tile_sum, i64 n_tiles)
{
    __shared__ double s_warp[FS_THREADS / 32];
    double carry = 0;
    for (i64 base = 0; base < n_tiles; base += FS_THREADS) {
        i64 i = base + threadIdx.x;
        double x = (i < n_tiles) ? tile_sum[i] : 0.0;
        double tot;
        double ex = block_excl_scan(x, s_warp, &tot);
        if (i < n_tiles) tile_sum[i] = carry + ex;
        carry += tot;
    }
}

__global__ void __launch_bounds__(FS_THREADS)
logfac_scan_apply(const i64 *__restrict__ counts, i64 n, const double *__restrict__ gtab,
                  const double *__restrict__ tile_prefix, double *__restrict__ out)
{
    __shared__ double s_warp[FS_THREADS / 32];
    const i64 base = (i64)blockIdx.x * FS_TILE + (i64)threadIdx.x * FS_ITEMS;
    double v[FS_ITEMS];
    double s = 0;
#pragma unroll
    for (int k = 0; k < FS_ITEMS; ++k) {
        double x = (base + k < n) ? __ldg(gtab + __ldg(counts + base + k) + 1) : 0.0;
        s += x;
        v[k] = s;
    }
    double tot;
    double off = tile_prefix[blockIdx.x] + block_excl_scan(s, s_warp, &tot);
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0.0;
#pragma unroll
    for (int k = 0; k < FS_ITEMS; ++k)
        if (base + k < n) out[base + k + 1] = off + v[k];
}

}  // namespace

int launch_scan_prepare(pasio_ctx *ctx, i64 *n_tiles, i64 *tile_elems)
{
    const i64 n = ctx->n;
    const i64 tiles = (n + 1 + SCAN_TILE - 1) / SCAN_TILE;    // positions 0..n
    PASIO_TRY(pasio_reserve(ctx, ctx->tilestate, (size_t)(tiles + 1) * 8 + 16));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->tilestate.p, 0, (size_t)(tiles + 1) * 8 + 16, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->scalars.p, 0, 16 * sizeof(i64), ctx->stream));
    ctx->scan_tiles_done = 0;
    if (n_tiles) *n_tiles = tiles;
    if (tile_elems) *tile_elems = SCAN_TILE;
    return PASIO_OK;
}

// The next `tiles` tiles: CTAs take consecutive tile numbers from the counter in the tile-state buffer, so
// successive launches continue where the previous one stopped and look back into its finished tiles.
int launch_scan_tiles(pasio_ctx *ctx, i64 tiles)
{
    if (tiles <= 0) return PASIO_OK;
    u64 *state = ctx->tilestate.as<u64>() + 2;
    unsigned *counter = ctx->tilestate.as<unsigned>();
    static int per_sm = 0;
    if (!per_sm) {
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_counts_kernel, SCAN_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
    }
    i64 grid = (i64)ctx->sm_count * per_sm;
    if (grid > tiles) grid = tiles;
    {
        TimingScope ts(ctx, TF_SCAN);
        CUDA_TRY(ctx, cudaMemsetAsync(counter, 0, 4, ctx->stream));          // per-launch tile counter
        scan_counts_kernel<<<(unsigned)grid, SCAN_THREADS, 0, ctx->stream>>>(
            ctx->counts.as<i64>(), ctx->n, ctx->cg.as<i64>(), ctx->cpbits.as<u64>(), state, counter, ctx->scan_tiles_done, tiles,
            ctx->scalars.as<i64>());
    }
    ctx->scan_tiles_done += tiles;
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_scan_counts(pasio_ctx *ctx)
{
    i64 tiles = 0;
    PASIO_TRY(launch_scan_prepare(ctx, &tiles, nullptr));
    return launch_scan_tiles(ctx, tiles);
}

int launch_expand_rle(pasio_ctx *ctx, const i64 *d_starts, const i64 *d_values, i64 n_runs)
{
    const i64 n = ctx->n;
    const i64 threads = (n + 7) / 8;
    TimingScope ts(ctx, TF_SCAN);
    expand_rle_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(
        d_starts, d_values, n_runs, n, ctx->counts.as<i64>());
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}

int launch_logfac_scan(pasio_ctx *ctx, double *d_out, cudaStream_t stream)
{
    const i64 n = ctx->n;
    if (!stream) stream = ctx->stream;
    // the lgamma table must cover max(counts)+1 (the scan kernel recorded the maximum)
    const i64 max_count = ctx->max_count;
    if (max_count + 2 > ctx->ntab[PASIO_TAB_LGAMMA]) {
        ctx->need[PASIO_TAB_LGAMMA] = max_count + 2;
        return pasio_fail(ctx, PASIO_E_TABLE_TOO_SHORT, "lgamma table has %lld entries, logfac needs %lld",
                          (long long)ctx->ntab[PASIO_TAB_LGAMMA], (long long)(max_count + 2));
    }
    const i64 tiles = (n + FS_TILE - 1) / FS_TILE;
    PASIO_TRY(pasio_reserve(ctx, ctx->fscan, (size_t)tiles * 8));
    const double *gtab = ctx->tab[PASIO_TAB_LGAMMA].as<double>();
    TimingScope ts(ctx, TF_SCORE, 3, stream);
    logfac_tile_sums<<<(unsigned)tiles, FS_THREADS, 0, stream>>>(ctx->counts.as<i64>(), n, gtab,
                                                                ctx->fscan.as<double>());
    logfac_scan_tile_sums<<<1, FS_THREADS, 0, stream>>>(ctx->fscan.as<double>(), tiles);
    logfac_scan_apply<<<(unsigned)tiles, FS_THREADS, 0, stream>>>(ctx->counts.as<i64>(), n, gtab,
                                                                 ctx->fscan.as<double>(), d_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return PASIO_OK;
}
