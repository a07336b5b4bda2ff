// Host-side text I/O of the bedgraph boundary, in C++ because the reference's per-line Python loop
// (6e6 nt/s) would dominate end-to-end time once the segmentation runs on the GPU.
//
//   pasio_bedgraph_parse   replaces BedgraphInterval.from_string / each_in_stream
//                          (/root/reference/src/pasio/dto/intervals.py:16-39): whitespace-separated
//                          chrom start stop count; blank lines skipped; a count that is not an
//                          integer literal is read as a float and truncated (int(float(x))).
//   pasio_format_segments  replaces the '%s\t%d\t%d\t%f\n' style writes of split_bedgraph_stream
//                          (/root/reference/src/pasio/process_bedgraph.py:71-89); snprintf("%f") rounds
//                          exactly like Python's '%f'.
// Both split their input into pieces (at line / segment boundaries) over a few host threads; the result is the
// same bytes / arrays the single-threaded walk gives.
// Grouping into contigs, gap filling and --split-at-gaps stay in Python (process_bedgraph.py), vectorised.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pasio_b200.h"

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Python int(): optional sign, decimal digits (no underscores handled: bedgraph has none)
inline bool parse_int(const char *s, const char *e, int64_t *out)
{
    if (s == e) return false;
    bool neg = false;
    if (*s == '+' || *s == '-') { neg = *s == '-'; ++s; }
    if (s == e) return false;
    uint64_t v = 0;
    for (; s < e; ++s) {
        if (*s < '0' || *s > '9') return false;
        const uint64_t d = (uint64_t)(*s - '0');
        if (v > (UINT64_C(9223372036854775807) - d) / 10) return false;   // beyond int64: not representable here (Python ints
        v = v * 10 + d;                                                    // are unbounded; numpy's int64 profile is not)
    }
    *out = neg ? -(int64_t)v : (int64_t)v;
    return true;
}

}  // namespace

// Counts the lines of a text buffer (an upper bound for the number of intervals).
extern "C" int64_t pasio_bedgraph_count_lines(const char *buf, int64_t len)
{
    int64_t n = 0;
    const char *p = buf, *end = buf + len;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        ++n;
        if (!nl) break;
        p = nl + 1;
    }
    return n;
}

namespace {

struct ParseOut {
    int64_t *starts, *stops, *counts, *name_off;
    int32_t *name_len;
    uint8_t *new_chrom;
};

// Parses the lines that start in [p, end) (p is a line start) into out[0 ..], at most cap of them.
// Returns true, or false with *err_line = number of the offending line counted from p.
bool parse_range(const char *buf, const char *p, const char *end, int64_t cap, const ParseOut &o, int64_t *n_out,
                 int64_t *nfloat_out, int64_t *err_line)
{
    int64_t n = 0, line_no = 0, nfloat = 0;
    const char *prev_name = nullptr;
    int32_t prev_len = 0;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const char *q = p;
        const char *tok[4], *tok_end[4];
        int nt = 0;
        while (q < le && nt < 4) {
            while (q < le && is_space(*q)) ++q;
            if (q >= le) break;
            tok[nt] = q;
            while (q < le && !is_space(*q)) ++q;
            tok_end[nt] = q;
            ++nt;
        }
        if (nt != 0) {                                   // blank lines are skipped
            if (nt < 4 || n >= cap) { *err_line = line_no; return false; }
            int64_t a, b, c;
            if (!parse_int(tok[1], tok_end[1], &a) || !parse_int(tok[2], tok_end[2], &b)) { *err_line = line_no; return false; }
            if (!parse_int(tok[3], tok_end[3], &c)) {
                char tmp[64];
                const size_t l = (size_t)(tok_end[3] - tok[3]);
                if (l >= sizeof tmp) { *err_line = line_no; return false; }
                memcpy(tmp, tok[3], l);
                tmp[l] = 0;
                // Python's float() takes no hexadecimal literals (strtod does)
                if (memchr(tmp, 'x', l) || memchr(tmp, 'X', l) || memchr(tmp, 'p', l) || memchr(tmp, 'P', l)) { *err_line = line_no; return false; }
                char *endp = nullptr;
                errno = 0;
                const double d = strtod(tmp, &endp);
                if (endp == tmp || *endp != 0 || d != d || d > 9.2e18 || d < -9.2e18) { *err_line = line_no; return false; }
                c = (int64_t)d;                              // int(float(x)) truncates toward zero
                ++nfloat;
            }
            o.starts[n] = a;
            o.stops[n] = b;
            o.counts[n] = c;
            o.name_off[n] = tok[0] - buf;
            o.name_len[n] = (int32_t)(tok_end[0] - tok[0]);
            o.new_chrom[n] = (prev_name == nullptr || prev_len != o.name_len[n] || memcmp(prev_name, tok[0], (size_t)prev_len) != 0);
            prev_name = tok[0];
            prev_len = o.name_len[n];
            ++n;
        }
        ++line_no;
        if (!nl) break;
        p = nl + 1;
    }
    *n_out = n;
    *nfloat_out = nfloat;
    return true;
}

int host_threads()
{
    static const int env = getenv("PASIO_B200_HOST_THREADS") ? atoi(getenv("PASIO_B200_HOST_THREADS")) : 0;
    if (env > 0) return env;
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(8u, hc ? hc / 2 : 4u));
}

}  // namespace

// Parses up to `cap` intervals.  Per interval: start, stop, count, and the byte range of the chromosome
// token inside buf (name_off, name_len).  new_chrom[i] = 1 when the token differs from the previous
// interval's (itertools.groupby over consecutive lines, process_bedgraph.py:33).
// float_counts receives how many counts needed the int(float(x)) conversion (the reference warns).
// Returns PASIO_OK, or PASIO_E_ARG with *n_out = index of the offending line.
// Large buffers are cut at line starts into one piece per host thread; every piece is parsed into the slot its
// line count reserves, then the pieces are closed up (blank lines leave holes) and new_chrom is fixed at the seams.
extern "C" int pasio_bedgraph_parse(const char *buf, int64_t len, int64_t cap, int64_t *starts, int64_t *stops,
                                    int64_t *counts, int64_t *name_off, int32_t *name_len, uint8_t *new_chrom,
                                    int64_t *n_out, int64_t *float_counts)
{
    const ParseOut all = {starts, stops, counts, name_off, name_len, new_chrom};
    const int T = len >= ((int64_t)4 << 20) ? host_threads() : 1;
    if (T > 1) {
        // piece boundaries at line starts
        std::vector<const char *> cut((size_t)T + 1);
        const char *end = buf + len;
        cut[0] = buf;
        cut[(size_t)T] = end;
        for (int t = 1; t < T; ++t) {
            const char *g = buf + len / T * t;
            if (g < cut[(size_t)t - 1]) g = cut[(size_t)t - 1];
            const char *nl = (const char *)memchr(g, '\n', (size_t)(end - g));
            cut[(size_t)t] = nl ? nl + 1 : end;
        }
        std::vector<int64_t> lines((size_t)T, 0), got((size_t)T, 0), nfl((size_t)T, 0), err((size_t)T, -1);
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < T; ++t)
                pool.emplace_back([&, t] { lines[(size_t)t] = pasio_bedgraph_count_lines(cut[(size_t)t], cut[(size_t)t + 1] - cut[(size_t)t]); });
            for (auto &th : pool) th.join();
        }
        int64_t total_lines = 0;
        for (int t = 0; t < T; ++t) total_lines += lines[(size_t)t];
        if (total_lines <= cap) {
            std::vector<int64_t> slot((size_t)T, 0);
            for (int t = 1; t < T; ++t) slot[(size_t)t] = slot[(size_t)t - 1] + lines[(size_t)t - 1];
            {
                std::vector<std::thread> pool;
                for (int t = 0; t < T; ++t)
                    pool.emplace_back([&, t] {
                        const int64_t s0 = slot[(size_t)t];
                        const ParseOut o = {starts + s0, stops + s0, counts + s0, name_off + s0, name_len + s0, new_chrom + s0};
                        int64_t e = 0;
                        if (!parse_range(buf, cut[(size_t)t], cut[(size_t)t + 1], lines[(size_t)t], o, &got[(size_t)t], &nfl[(size_t)t], &e))
                            err[(size_t)t] = e;
                    });
                for (auto &th : pool) th.join();
            }
            int64_t before = 0;
            for (int t = 0; t < T; ++t) {
                if (err[(size_t)t] >= 0) { *n_out = before + err[(size_t)t]; return PASIO_E_ARG; }
                before += lines[(size_t)t];
            }
            int64_t n = 0, nfloat = 0;
            for (int t = 0; t < T; ++t) {
                const int64_t s0 = slot[(size_t)t], g = got[(size_t)t];
                if (g > 0) {
                    if (s0 != n) {
                        memmove(starts + n, starts + s0, (size_t)g * 8);
                        memmove(stops + n, stops + s0, (size_t)g * 8);
                        memmove(counts + n, counts + s0, (size_t)g * 8);
                        memmove(name_off + n, name_off + s0, (size_t)g * 8);
                        memmove(name_len + n, name_len + s0, (size_t)g * 4);
                        memmove(new_chrom + n, new_chrom + s0, (size_t)g);
                    }
                    if (n > 0)          // the piece's first interval against the one before the seam
                        new_chrom[n] = (name_len[n] != name_len[n - 1] ||
                                        memcmp(buf + name_off[n], buf + name_off[n - 1], (size_t)name_len[n]) != 0);
                }
                n += g;
                nfloat += nfl[(size_t)t];
            }
            *n_out = n;
            if (float_counts) *float_counts = nfloat;
            return PASIO_OK;
        }
    }
    int64_t n = 0, nfloat = 0, e = 0;
    if (!parse_range(buf, buf, buf + len, cap, all, &n, &nfloat, &e)) { *n_out = e; return PASIO_E_ARG; }
    *n_out = n;
    if (float_counts) *float_counts = nfloat;
    return PASIO_OK;
}

namespace {

// '%f' of a double and '%d' of an integer with std::to_chars: by the standard the fixed/6 conversion writes the
// characters printf("%f") writes in the C locale (checked on 2e7 random doubles), several times faster.
inline char *put_f(char *p, char *end, double v)
{
    if (std::isnan(v)) { memcpy(p, "nan", 3); return p + 3; }          // Python writes 'nan' whatever the sign bit
    if (!std::isfinite(v)) return p + snprintf(p, (size_t)(end - p), "%f", v);
    return std::to_chars(p, end, v, std::chars_format::fixed, 6).ptr;
}
inline char *put_d(char *p, char *end, long long v) { return std::to_chars(p, end, v).ptr; }

// lines of segments [k0, k1) appended to dst
void format_range(const char *chrom, size_t clen, int64_t offset, const int64_t *splits, int64_t k0, int64_t k1,
                  const double *means, const double *lmm, int mode, std::string &dst)
{
    char line[800];
    char *const end = line + sizeof line - 8;       // two int64, one more, two %f of doubles up to 1e308: < 700 chars; room for the separators
    dst.reserve((size_t)(k1 - k0) * (clen + 40));
    for (int64_t k = k0; k < k1; ++k) {
        const long long a = (long long)(splits[k] + offset), b = (long long)(splits[k + 1] + offset);
        char *p = line;
        *p++ = '\t';
        p = put_d(p, end, a);
        *p++ = '\t';
        p = put_d(p, end, b);
        if (mode != 1) {
            *p++ = '\t';
            p = put_f(p, end, means[k]);
        }
        if (mode == 2) {
            *p++ = '\t';
            p = put_d(p, end, b - a);
            *p++ = '\t';
            p = put_f(p, end, lmm[k]);
        }
        *p++ = '\n';
        dst.append(chrom, clen);
        dst.append(line, (size_t)(p - line));
    }
}

}  // namespace

// mode 0: chrom start stop mean ; 1: chrom start stop ; 2: chrom start stop mean length lmm
// Writes at most cap bytes; returns the number of bytes written, or -(bytes needed) if cap is too small.
extern "C" int64_t pasio_format_segments(const char *chrom, int64_t offset, const int64_t *splits, int64_t n_splits,
                                         const double *means, const double *lmm, int mode, char *out, int64_t cap)
{
    const size_t clen = strlen(chrom);
    const int64_t nseg = n_splits > 0 ? n_splits - 1 : 0;
    const int T = nseg >= 100000 ? host_threads() : 1;
    std::vector<std::string> part((size_t)T);
    if (T == 1) {
        format_range(chrom, clen, offset, splits, 0, nseg, means, lmm, mode, part[0]);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t)
            pool.emplace_back([&, t] {
                format_range(chrom, clen, offset, splits, nseg * t / T, nseg * (t + 1) / T, means, lmm, mode, part[(size_t)t]);
            });
        for (auto &th : pool) th.join();
    }
    int64_t need = 0;
    for (auto &p : part) need += (int64_t)p.size();
    if (need > cap) return -need;
    int64_t w = 0;
    for (auto &p : part) {
        memcpy(out + w, p.data(), p.size());
        w += (int64_t)p.size();
    }
    return w;
}

// The same for a batch of contigs that were segmented as one super-contig (pasio_contig_load* with offsets): splits
// are positions in the super-contig and contain every contig boundary; first_split[c] is the index of contig c's
// first split point (first_split[n_contigs] = index of the last split point of the batch), shift[c] is added to its
// positions (chromosome start minus the contig's offset in the super-contig), names[name_off[c] .. name_off[c+1])
// is its name.  Lines come out in contig order, exactly those of per-contig pasio_format_segments calls.
extern "C" int64_t pasio_format_segments_batch(const char *names, const int64_t *name_off, const int64_t *shift,
                                               const int64_t *first_split, int64_t n_contigs, const int64_t *splits,
                                               const double *means, const double *lmm, int mode, char *out, int64_t cap)
{
    const int64_t nseg = n_contigs > 0 ? first_split[n_contigs] - first_split[0] : 0;
    const int64_t seg0 = n_contigs > 0 ? first_split[0] : 0;
    const int T = nseg >= 100000 ? host_threads() : 1;
    std::vector<std::string> part((size_t)T);
    auto work = [&](int t) {
        const int64_t k0 = seg0 + nseg * t / T, k1 = seg0 + nseg * (t + 1) / T;
        if (k0 >= k1) return;
        // contig of segment k0: last c with first_split[c] <= k0
        int64_t c = std::upper_bound(first_split, first_split + n_contigs + 1, k0) - first_split - 1;
        std::string &dst = part[(size_t)t];
        int64_t k = k0;
        while (k < k1) {
            const int64_t kend = std::min(k1, first_split[c + 1]);
            const std::string name(names + name_off[c], (size_t)(name_off[c + 1] - name_off[c]));
            format_range(name.c_str(), name.size(), shift[c], splits, k, kend, means, lmm, mode, dst);
            k = kend;
            ++c;
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    int64_t need = 0;
    for (auto &p : part) need += (int64_t)p.size();
    if (need > cap) return -need;
    int64_t w = 0;
    for (auto &p : part) {
        memcpy(out + w, p.data(), p.size());
        w += (int64_t)p.size();
    }
    return w;
}

// Groups the parsed intervals into contigs and turns every contig into run lengths / run values: the accumulation
// loop of parse_bedgraph_stream and interval_groups (/root/reference/src/pasio/process_bedgraph.py:26-60).
// Consecutive lines of one chromosome form a group; with split_at_gaps a group also ends where an interval does not
// start at the previous stop; otherwise a zero run is inserted between non-adjacent intervals (when the previous
// stop is non-zero, as in the reference's `if previous_stop and ...`).  Runs of non-positive length are dropped.
// Outputs: run_len / run_val (capacity 2n), and per group its first line and its first run (group_run has
// n_groups + 1 entries).  Returns the number of groups.
extern "C" int64_t pasio_bedgraph_runs(const int64_t *starts, const int64_t *stops, const int64_t *counts,
                                       const uint8_t *new_chrom, int64_t n, int split_at_gaps, int64_t *run_len,
                                       int64_t *run_val, int64_t *group_line, int64_t *group_run, int64_t *n_runs)
{
    int64_t g = 0, r = 0;
    for (int64_t i = 0; i < n; ++i) {
        const bool adjacent = i == 0 || starts[i] == stops[i - 1];
        const bool cut = new_chrom[i] || (split_at_gaps && !adjacent);
        if (cut) {
            group_line[g] = i;
            group_run[g] = r;
            ++g;
        } else if (!split_at_gaps && !adjacent && stops[i - 1] != 0) {
            const int64_t gap = starts[i] - stops[i - 1];
            if (gap > 0) { run_len[r] = gap; run_val[r] = 0; ++r; }
        }
        const int64_t len = stops[i] - starts[i];
        if (len > 0) { run_len[r] = len; run_val[r] = counts[i]; ++r; }
    }
    group_run[g] = r;
    *n_runs = r;
    return g;
}


// ---- int64 -> int32 packing for the narrowed upload (api.cu: NarrowUpload) ---------------------------------------
// dst[i] = (int32) src[i]; returns the OR of all values: bits 31..63 set <=> some count is negative or >= 2^31.
// Streaming stores: the packed slice goes to a page-locked buffer the DMA engine reads, not back to this core.
#include <immintrin.h>
__attribute__((target("avx2"))) static uint64_t narrow_slice_avx2(int32_t *dst, const int64_t *src, size_t n)
{
    __m256i acc = _mm256_setzero_si256();
    const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (; i + 8 <= n; i += 8) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 4));
            acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, idx)));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 4), _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, idx)));
        }
    }
    uint64_t o[4];
    _mm256_storeu_si256(reinterpret_cast<__m256i *>(o), acc);
    uint64_t r = o[0] | o[1] | o[2] | o[3];
    for (; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (int32_t)src[i]; }
    _mm_sfence();
    return r;
}

// the same to uint16 (coverage counts almost always fit): values >= 2^16 saturate, the returned OR tells the caller to redo the slice wider
__attribute__((target("avx2"))) static uint64_t narrow_slice16_avx2(uint16_t *dst, const int64_t *src, size_t n)
{
    __m256i acc = _mm256_setzero_si256();
    const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (; i + 8 <= n; i += 8) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 4));
            acc = _mm256_or_si256(acc, _mm256_or_si256(a, b));
            const __m128i lo = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, idx));
            const __m128i hi = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, idx));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_packus_epi32(lo, hi));
        }
    }
    uint64_t o[4];
    _mm256_storeu_si256(reinterpret_cast<__m256i *>(o), acc);
    uint64_t r = o[0] | o[1] | o[2] | o[3];
    for (; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (uint16_t)src[i]; }
    _mm_sfence();
    return r;
}

// and to uint8 (most coverage is below 256 everywhere)
__attribute__((target("avx2"))) static uint64_t narrow_slice8_avx2(uint8_t *dst, const int64_t *src, size_t n)
{
    __m256i acc = _mm256_setzero_si256();
    const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (; i + 16 <= n; i += 16) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 4));
            const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 8));
            const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 12));
            acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
            const __m128i ab = _mm_packus_epi32(_mm256_castsi256_si128(_mm256_permutevar8x32_epi32(a, idx)),
                                                _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(b, idx)));
            const __m128i cd = _mm_packus_epi32(_mm256_castsi256_si128(_mm256_permutevar8x32_epi32(c, idx)),
                                                _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(d, idx)));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), _mm_packus_epi16(ab, cd));
        }
    }
    uint64_t o[4];
    _mm256_storeu_si256(reinterpret_cast<__m256i *>(o), acc);
    uint64_t r = o[0] | o[1] | o[2] | o[3];
    for (; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (uint8_t)src[i]; }
    _mm_sfence();
    return r;
}

__attribute__((visibility("hidden"))) uint64_t pasio_narrow_slice8(uint8_t *dst, const int64_t *src, size_t n)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return narrow_slice8_avx2(dst, src, n);
    uint64_t r = 0;
    for (size_t i = 0; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (uint8_t)src[i]; }
    return r;
}

__attribute__((visibility("hidden"))) uint64_t pasio_narrow_slice16(uint16_t *dst, const int64_t *src, size_t n)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return narrow_slice16_avx2(dst, src, n);
    uint64_t r = 0;
    for (size_t i = 0; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (uint16_t)src[i]; }
    return r;
}

__attribute__((visibility("hidden"))) uint64_t pasio_narrow_slice(int32_t *dst, const int64_t *src, size_t n)
{
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) return narrow_slice_avx2(dst, src, n);
    uint64_t r = 0;
    for (size_t i = 0; i < n; ++i) { r |= (uint64_t)src[i]; dst[i] = (int32_t)src[i]; }
    return r;
}
