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

constexpr int SCAN_PER_THREAD = SCAN_TILE / SCAN_THREADS;          // 16 consecutive elements per thread
// element e of the tile lives at e + 2 * (e / 16) in shared memory: a thread's 16 elements are 8 aligned 16-byte pairs,
// and the 144-byte stride between threads spreads them over the banks
__device__ __forceinline__ int scan_slot(int e) { return e + 2 * (e >> 4); }

// The tile is read with fully coalesced 512-byte warp accesses, transposed through shared memory so that every thread
// owns 16 CONSECUTIVE elements (a local running sum, the change-point tests and the sign / maximum checks are then a
// dozen instructions per element instead of the per-slab warp scans of the first version, which made the kernel
// issue-bound: profiles/r02_window_dp_scan_ncu_full_bench_chr1.txt, 92 instructions per element, 49 % of the copy
// peak), scanned across threads once per tile, and written back the same coalesced way.  cg[p] is the EXCLUSIVE prefix
// at p, so the (cg[2m], cg[2m+1]) pair is an aligned 16-byte store.
__global__ void __launch_bounds__(SCAN_THREADS, 3)
scan_counts_kernel(const i64 *__restrict__ counts, i64 n, i64 *__restrict__ cg,
                   u64 *__restrict__ cpwords, u64 *tile_state, unsigned *tile_counter, i64 first_tile, i64 n_tiles,
                   i64 *scalars /* [0]=total, [1]=negative seen, [2]=largest count */)
{
    __shared__ __align__(16) i64 sbuf[SCAN_TILE + 2 * (SCAN_TILE / 16)];
    __shared__ unsigned s_tile;
    __shared__ i64 s_warp[SCAN_THREADS / 32];
    __shared__ i64 s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Persistent CTAs (one resident wave) take tile numbers from a per-launch counter: tiles start in issue order, so the
    // look-back never waits on an unscheduled tile.
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
        __syncthreads();
        if ((i64)s_tile >= n_tiles) break;
        const i64 tile = first_tile + s_tile;
        const i64 tbase = tile * SCAN_TILE;

        // ---- coalesced load -> shared memory (warp w: elements [512w, 512w + 512), 8 slabs of 32 lanes x one pair) ----
#pragma unroll
        for (int k = 0; k < SCAN_SLABS; ++k) {
            const int e = warp * SCAN_WARP_ELEMS + 64 * k + 2 * lane;
            const i64 g = tbase + e;
            longlong2 t = make_longlong2(0, 0);
            if (g + 1 < n) t = __ldg(reinterpret_cast<const longlong2 *>(counts + g));
            else if (g < n) t.x = __ldg(counts + g);
            *reinterpret_cast<longlong2 *>(sbuf + scan_slot(e)) = t;
        }
        const i64 before_tile = (tid == 0 && tbase > 0 && tbase - 1 < n) ? __ldg(counts + tbase - 1) : 0;
        __syncthreads();

        // ---- 16 consecutive elements per thread ----
        const int e0 = tid * SCAN_PER_THREAD;
        const i64 p0 = tbase + e0;                      // position of this thread's first element
        i64 v[SCAN_PER_THREAD];
#pragma unroll
        for (int q = 0; q < SCAN_PER_THREAD / 2; ++q) {
            const longlong2 t = *reinterpret_cast<const longlong2 *>(sbuf + scan_slot(e0) + 2 * q);
            v[2 * q] = t.x;
            v[2 * q + 1] = t.y;
        }
        i64 prev = __shfl_up_sync(0xffffffffu, v[SCAN_PER_THREAD - 1], 1);
        if (lane == 0) prev = tid == 0 ? before_tile : sbuf[scan_slot(e0 - 1)];
        i64 run = 0, any = 0, vmax = 0;
        unsigned bits = 0;
#pragma unroll
        for (int k = 0; k < SCAN_PER_THREAD; ++k) {
            const i64 p = p0 + k;
            // change point at p: counts[p-1] != counts[p], 1 <= p <= n-1 (constants_reducer.py:16-17)
            if (v[k] != prev && p >= 1 && p <= n - 1) bits |= 1u << k;
            prev = v[k];
            any |= v[k];
            vmax = max(vmax, v[k]);
            const i64 x = v[k];
            v[k] = run;                                 // exclusive prefix inside the thread
            run += x;
        }
        if (any < 0) scalars[1] = 1;                    // a negative count sets the sign bit of the OR
        // largest count (the log-factorial table must reach it): one atomic per warp, and only when it raises the maximum
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
        if (lane == 0 && vmax > *reinterpret_cast<volatile i64 *>(scalars + 2)) atomicMax(reinterpret_cast<long long *>(scalars + 2), vmax);
        // change-point words: four threads per 64-bit word
        {
            u64 w = (u64)bits << (16 * (lane & 3));
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            if ((lane & 3) == 0 && p0 <= n) cpwords[p0 >> 6] = w;
        }
        // scan of the thread totals over the warp, then over the CTA
        i64 incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const i64 o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
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
        const i64 off = s_prefix + warp_off + incl - run;
        // prefix sums back to shared memory (same slots: every thread rewrites its own 16), then coalesced stores
#pragma unroll
        for (int q = 0; q < SCAN_PER_THREAD / 2; ++q)
            *reinterpret_cast<longlong2 *>(sbuf + scan_slot(e0) + 2 * q) = make_longlong2(off + v[2 * q], off + v[2 * q + 1]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SCAN_SLABS; ++k) {
            const int e = warp * SCAN_WARP_ELEMS + 64 * k + 2 * lane;
            const i64 g = tbase + e;
            const longlong2 t = *reinterpret_cast<const longlong2 *>(sbuf + scan_slot(e));
            if (g + 1 <= n) {
                *reinterpret_cast<longlong2 *>(cg + g) = t;
                if (g + 1 == n) scalars[0] = t.y;
            } else if (g <= n) {
                cg[g] = t.x;
                if (g == n) scalars[0] = t.x;
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
