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
// Grouping into contigs, gap filling and --split-at-gaps stay in Python (process_bedgraph.py), vectorised.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
    int64_t v = 0;
    for (; s < e; ++s) {
        if (*s < '0' || *s > '9') return false;
        v = v * 10 + (*s - '0');
    }
    *out = neg ? -v : v;
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

// Parses up to `cap` intervals.  Per interval: start, stop, count, and the byte range of the chromosome
// token inside buf (name_off, name_len).  new_chrom[i] = 1 when the token differs from the previous
// interval's (itertools.groupby over consecutive lines, process_bedgraph.py:33).
// float_counts receives how many counts needed the int(float(x)) conversion (the reference warns).
// Returns PASIO_OK, or PASIO_E_ARG with *n_out = index of the offending line.
extern "C" int pasio_bedgraph_parse(const char *buf, int64_t len, int64_t cap, int64_t *starts, int64_t *stops,
                                    int64_t *counts, int64_t *name_off, int32_t *name_len, uint8_t *new_chrom,
                                    int64_t *n_out, int64_t *float_counts)
{
    const char *p = buf, *end = buf + len;
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
            if (nt < 4 || n >= cap) { *n_out = line_no; return PASIO_E_ARG; }
            int64_t a, b, c;
            if (!parse_int(tok[1], tok_end[1], &a) || !parse_int(tok[2], tok_end[2], &b)) { *n_out = line_no; return PASIO_E_ARG; }
            if (!parse_int(tok[3], tok_end[3], &c)) {
                char tmp[64];
                const size_t l = (size_t)(tok_end[3] - tok[3]);
                if (l >= sizeof tmp) { *n_out = line_no; return PASIO_E_ARG; }
                memcpy(tmp, tok[3], l);
                tmp[l] = 0;
                char *endp = nullptr;
                errno = 0;
                const double d = strtod(tmp, &endp);
                if (endp == tmp || *endp != 0 || d != d || d > 9.2e18 || d < -9.2e18) { *n_out = line_no; return PASIO_E_ARG; }
                c = (int64_t)d;                              // int(float(x)) truncates toward zero
                ++nfloat;
            }
            starts[n] = a;
            stops[n] = b;
            counts[n] = c;
            name_off[n] = tok[0] - buf;
            name_len[n] = (int32_t)(tok_end[0] - tok[0]);
            new_chrom[n] = (prev_name == nullptr || prev_len != name_len[n] || memcmp(prev_name, tok[0], (size_t)prev_len) != 0);
            prev_name = tok[0];
            prev_len = name_len[n];
            ++n;
        }
        ++line_no;
        if (!nl) break;
        p = nl + 1;
    }
    *n_out = n;
    if (float_counts) *float_counts = nfloat;
    return PASIO_OK;
}

// mode 0: chrom start stop mean ; 1: chrom start stop ; 2: chrom start stop mean length lmm
// Writes at most cap bytes; returns the number of bytes written, or -(bytes needed estimate) if cap is too small.
extern "C" int64_t pasio_format_segments(const char *chrom, int64_t offset, const int64_t *splits, int64_t n_splits,
                                         const double *means, const double *lmm, int mode, char *out, int64_t cap)
{
    const size_t clen = strlen(chrom);
    int64_t w = 0;
    for (int64_t k = 0; k + 1 < n_splits; ++k) {
        if (cap - w < (int64_t)clen + 400) return -((int64_t)(clen + 400) * (n_splits - 1));
        memcpy(out + w, chrom, clen);
        w += (int64_t)clen;
        const long long a = (long long)(splits[k] + offset), b = (long long)(splits[k + 1] + offset);
        int m;
        if (mode == 1) m = snprintf(out + w, 400, "\t%lld\t%lld\n", a, b);
        else if (mode == 0) m = snprintf(out + w, 400, "\t%lld\t%lld\t%f\n", a, b, means[k]);
        else m = snprintf(out + w, 400, "\t%lld\t%lld\t%f\t%lld\t%f\n", a, b, means[k], b - a, lmm[k]);
        if (m < 0 || m >= 400) return -((int64_t)(clen + 400) * (n_splits - 1));
        w += m;
    }
    return w;
}
