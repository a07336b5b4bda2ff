"""CPU oracle for the Pasio segmentation hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (autosome-ru/pasio
v1.1.3, pure Python + numpy/scipy).  It exists so that the CUDA path can be
checked on a box where /root/reference does not exist.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it; the product package pasio_b200 never does.

Parity pinning: tests/golden/*.npz were produced by oracle/make_golden.py from the
UNMODIFIED reference imported from /root/reference/src (with the future shim in
oracle/future_shim); tests/test_oracle_golden.py checks this restatement against
them bit for bit (splits, scores).  The arithmetic itself lives in numpy.log /
scipy.special.gammaln (un-pinned third-party; numpy 2.3.5 / scipy 1.18.1 here),
called exactly where the reference calls them, so on one host the oracle and the
reference see identical table values.

Each function cites the reference file:line it follows (paths under
/root/reference/src/pasio/).
"""
import numpy as np
import scipy.special

CACHE_SIZE = 1 << 20  # cached_log.py:6,33 default cache_size


class Tables(object):
    """log / log-gamma look-up with the reference's table-or-direct split.

    cached_log.py:5-29 (LogComputer) and :32-56 (LogGammaComputer);
    instantiated as in log_marginal_likelyhood.py:14-16.
    """

    def __init__(self, alpha, beta, cache_size=CACHE_SIZE):
        self.alpha = alpha
        self.beta = beta
        self.cache_size = cache_size
        k = np.arange(cache_size)
        self.log_tab = np.log(k + beta)                      # cached_log.py:9
        self.lgam_tab = scipy.special.gammaln(k + 0)         # cached_log.py:36, shift=0
        self.lgam_alpha_tab = scipy.special.gammaln(k + alpha)  # cached_log.py:36, shift=alpha

    def _lookup(self, table, direct, x):
        # cached_log.py:24-29 / :51-56 (compute_for_array_unbound)
        x = np.asarray(x)
        out = np.zeros(x.shape)
        small = x < self.cache_size
        out[small] = table[x[small]]
        out[~small] = direct(x[~small])
        return out

    def log(self, x):
        return self._lookup(self.log_tab, lambda v: np.log(v + self.beta), x)

    def lgam(self, x):
        return self._lookup(self.lgam_tab, lambda v: scipy.special.gammaln(v + 0), x)

    def lgam_alpha(self, x):
        return self._lookup(self.lgam_alpha_tab, lambda v: scipy.special.gammaln(v + self.alpha), x)

    def log_number(self, x):
        # cached_log.py:11-15
        return self.log_tab[x] if x < self.cache_size else np.log(x + self.beta)

    def lgam_alpha_number(self, x):
        # cached_log.py:38-42
        return self.lgam_alpha_tab[x] if x < self.cache_size else scipy.special.gammaln(x + self.alpha)


def normalise_alpha(alpha):
    """log_marginal_likelyhood.py:9-12: integral floats become int."""
    if isinstance(alpha, float) and alpha.is_integer():
        return int(alpha)
    return alpha


class Scorer(object):
    """log_marginal_likelyhood.py:45-132 (base + Int/Real alpha variants)."""

    def __init__(self, counts, cands, tables):
        counts = np.asarray(counts)
        cands = np.asarray(cands)
        # log_marginal_likelyhood.py:30-40
        assert counts.dtype == int and np.all(counts >= 0) and len(counts) > 0
        assert cands[0] == 0 and cands[-1] == len(counts) and np.all(cands[1:] > cands[:-1])
        self.t = tables
        self.alpha = tables.alpha
        self.int_alpha = isinstance(self.alpha, (int, np.integer))
        self.cands = cands
        self.cumsum = np.hstack([0, np.cumsum(counts)])[cands]                    # :57
        logfac = tables.lgam(counts + 1)                                          # :59
        self.logfac_cumsum = np.hstack([0, np.cumsum(logfac)])[cands]             # :60
        self.pen = self.alpha * tables.log_number(0) - tables.lgam_alpha_number(0)  # :62

    def row(self, stop):
        """all_suffixes_self_score: :105-115 (int alpha) / :121-132 (real alpha)."""
        shifted = (self.alpha + self.cumsum[stop]) - self.cumsum[0:stop]
        lengths = self.cands[stop] - self.cands[:stop]
        if self.int_alpha:
            add = self.t.lgam(shifted)
        else:
            add = self.t.lgam_alpha(self.cumsum[stop] - self.cumsum[0:stop])
        sub = shifted * self.t.log(lengths)
        return add - sub

    def scores(self):
        # :67-74
        seg_len = np.diff(self.cands)
        seg_cnt = np.diff(self.cumsum)
        add = self.t.lgam_alpha(seg_cnt)
        sub = (seg_cnt + self.alpha) * self.t.log(seg_len)
        return (add - sub) + self.pen

    def mean_counts(self):
        # :80-83
        return np.diff(self.cumsum) / np.diff(self.cands)

    def log_marginal_likelyhoods(self):
        # :76-78
        return self.scores() - np.diff(self.logfac_cumsum)

    def total_sum_logfac(self):
        # :64-65
        return self.logfac_cumsum[-1]


def backtrace(prev):
    """square_splitter.py:102-109 (collect_split_points)."""
    k = len(prev) - 1
    chain = [k]
    while k != 0:
        k = int(prev[k])
        chain.append(k)
    return chain[::-1]


def square_split(counts, cands, tables):
    """square_splitter.py:67-100 (split_without_normalizations).

    Returns (score, split_positions, prefix_scores, previous_splits)."""
    sc = Scorer(counts, cands, tables)
    n = len(cands)
    prefix = np.empty(n)
    prefix[0] = 0
    prev = np.empty(n, dtype=int)
    prev[0] = 0
    for j in range(1, n):
        t = sc.row(j)
        t += prefix[:j]                      # :88
        k = np.argmax(t)                     # :90 first index among maxima
        prev[j] = k
        prefix[j] = t[k] + sc.pen            # :94
    idx = backtrace(prev)
    return prefix[-1], np.asarray(cands)[idx], prefix, prev


def square_split_regularized(counts, cands, tables, len_mult=0, len_fn=lambda x: x,
                             num_mult=0, num_fn=lambda x: x):
    """square_splitter.py:29-65 (split_with_normalizations)."""
    sc = Scorer(counts, cands, tables)
    cands = np.asarray(cands)
    n = len(cands)
    prefix = np.empty(n)
    prefix[0] = 0
    prev = np.empty(n, dtype=int)
    prev[0] = 0
    nsplits = np.zeros(n)
    for j in range(1, n):
        t = sc.row(j)
        t += prefix[:j]
        if num_mult != 0:
            t -= num_mult * num_fn(nsplits[:j] + 1)
            t[0] += num_mult * num_fn(1)
        if len_mult != 0:
            t -= (len_mult * len_fn(cands[j] - cands[:j]))[:j]
        k = np.argmax(t)
        prev[j] = k
        if k != 0:
            nsplits[j] = nsplits[k] + 1
        prefix[j] = t[k] + sc.pen
    idx = backtrace(prev)
    return prefix[-1], cands[idx]


def not_zero(counts, cands):
    """constants_reducer.py:5-11."""
    if np.all(counts == 0):
        return np.array([0, len(counts)])
    return cands


def not_constant(counts, cands):
    """constants_reducer.py:14-21."""
    (left,) = np.where(counts[:-1] != counts[1:])
    change = 1 + left
    keep = np.intersect1d(cands, change, assume_unique=True)
    return np.hstack([0, keep, len(counts)])


def window_ranges(m, window_size, window_shift):
    """dto/sliding_window.py:9-15 -- index ranges [start, stop) over m candidates."""
    return [(st, min(st + window_size + 1, m)) for st in range(0, m - 1, window_shift)]


def _base_reduce(counts, cands, tables, constraint):
    # default_splitters.py:52-59 base_splitter graph, reducer_combiner.py:5-8
    if constraint == 'constants':
        cands = not_constant(counts, cands)
    elif constraint == 'zeros':
        cands = not_zero(counts, cands)
    elif constraint != 'none':
        raise ValueError(constraint)
    return square_split(counts, cands, tables)[1]       # square_splitter.py:19-21


def sliding_window_round(counts, cands, tables, window_size, window_shift, constraint):
    """sliding_window_reducer.py:10-29 -- one round, object-faithful (slices + re-basing)."""
    keep = set([0, len(counts)])                                         # :22
    for st, en in window_ranges(len(cands), window_size, window_shift):  # :23
        win = cands[st:en]
        a, b = win[0], win[-1]                                           # :11-12
        reduced = _base_reduce(counts[a:b], win - a, tables, constraint) + a  # :15-18
        keep.update(reduced.tolist())                                    # :25
    return np.array(sorted(keep))                                        # :29


def round_reduce(counts, cands, tables, window_size, window_shift, constraint, num_rounds=None):
    """round_reducer.py:10-31.  Returns (candidates, per-round sizes)."""
    rounds = len(counts) if num_rounds is None else num_rounds
    rounds = max(1, rounds)
    sizes = [len(cands)]
    for _ in range(rounds):
        new = sliding_window_round(counts, cands, tables, window_size, window_shift, constraint)
        if np.array_equal(new, cands):
            return new, sizes
        assert len(new) < len(cands)
        cands = new
        sizes.append(len(cands))
    return cands, sizes


def nop_split(counts, cands, tables):
    """nop_splitter.py:15-18."""
    return np.sum(Scorer(counts, cands, tables).scores()), cands


def default_pipeline(counts, tables, window_size=2500, window_shift=1250,
                     constraint='constants', num_rounds=None):
    """default_splitters.py:65-66 graph driven as in segmentation.py:5-20.

    Returns dict(score, splits, mean_counts, lmm, sum_logfac, sizes)."""
    counts = np.asarray(counts)
    cands = np.arange(len(counts) + 1)                                    # segmentation.py:8
    cands, sizes = round_reduce(counts, cands, tables, window_size, window_shift,
                                constraint, num_rounds)
    score, splits = nop_split(counts, cands, tables)
    sc = Scorer(counts, splits, tables)                                   # segmentation.py:10
    return dict(score=score, splits=splits, mean_counts=sc.mean_counts(),
                lmm=sc.log_marginal_likelyhoods(), sum_logfac=sc.total_sum_logfac(),
                sizes=np.array(sizes))


def exact_pipeline(counts, tables):
    """--algorithm exact (default_splitters.py:48-49) driven as in segmentation.py:5-20."""
    counts = np.asarray(counts)
    score, splits, _, _ = square_split(counts, np.arange(len(counts) + 1), tables)
    sc = Scorer(counts, splits, tables)
    return dict(score=score, splits=splits, mean_counts=sc.mean_counts(),
                lmm=sc.log_marginal_likelyhoods(), sum_logfac=sc.total_sum_logfac())


# ---------------------------------------------------------------------------
# Flat restatement of a sliding-window round: no slicing, no per-window
# cumsum.  Verified equal to sliding_window_round() in tests/test_oracle_golden.py;
# used (through oracle/dp_oracle.c) to check the CUDA path at sizes where the
# object-faithful loop above would take minutes.
# ---------------------------------------------------------------------------

def extended_tables(tables, n_log, n_gam):
    """Host tables extended past 2**20 by the same numpy/scipy calls (SURVEY 7.3:
    bit-identical to the reference's unbound path).  Returns (Lg, G, Ga)."""
    n_log = max(int(n_log), 2)
    n_gam = max(int(n_gam), 2)
    lg = np.log(np.arange(n_log) + tables.beta)
    g = scipy.special.gammaln(np.arange(n_gam) + 0)
    ga = scipy.special.gammaln(np.arange(n_gam) + tables.alpha)
    return lg, g, ga


# ---------------------------------------------------------------------------
# Bedgraph text -> per-contig dense profiles, the reference's way (per-line Python loop).
# dto/intervals.py:16-39 (from_string / each_in_stream), process_bedgraph.py:9-60
# (fill_interval_gaps, interval_groups, parse_bedgraph_stream).
# ---------------------------------------------------------------------------

def parse_bedgraph_text(text, split_at_gaps=False):
    """-> list of (chrom, dense int64 profile, chrom_start)"""
    intervals = []
    for line in text.splitlines():
        line = line.strip()
        if line == '':
            continue
        chrom, start, stop, count_str = line.split()[0:4]          # intervals.py:17
        try:
            count = int(count_str)
        except ValueError:
            count = int(float(count_str))                          # intervals.py:23
        intervals.append((chrom, int(start), int(stop), count))
    # consecutive grouping by chromosome (itertools.groupby, process_bedgraph.py:33)
    groups, k = [], 0
    while k < len(intervals):
        j = k
        while j < len(intervals) and intervals[j][0] == intervals[k][0]:
            j += 1
        groups.append(intervals[k:j])
        k = j
    contigs = []
    for grp in groups:
        if split_at_gaps:                                          # slice_when(intervals_not_adjacent), :34-36
            parts, cur = [], [grp[0]]
            for prev, nxt in zip(grp[:-1], grp[1:]):
                if prev[2] != nxt[1]:
                    parts.append(cur)
                    cur = []
                cur.append(nxt)
            parts.append(cur)
        else:                                                      # fill_interval_gaps, :9-16
            filled, previous_stop = [], None
            for iv in grp:
                if previous_stop and previous_stop != iv[1]:
                    filled.append((iv[0], previous_stop, iv[1], 0))
                filled.append(iv)
                previous_stop = iv[2]
            parts = [filled]
        for part in parts:                                         # :49-60
            data = []
            for (chrom, start, stop, cov) in part:
                data.extend([cov] * (stop - start))
            contigs.append((part[0][0], np.array(data, dtype=int), part[0][1]))
    return contigs


def split_bedgraph_text(text, tables, window_size=2500, window_shift=1250, constraint='constants', num_rounds=None,
                        split_at_gaps=False, output_mode='bedgraph', threads=1):
    """process_bedgraph.py:67-92 with the default splitter graph: bedgraph text in, the reference's output text out.
    The rounds run through the flat C restatement (oracle/c_oracle.py) so that hundreds of contigs take seconds;
    scoring and the '%' formatting are the reference's (process_bedgraph.py:71-89, log_marginal_likelyhood.py:67-83)."""
    from . import c_oracle
    out = []
    for chrom, counts, chrom_start in parse_bedgraph_text(text, split_at_gaps):
        fo = c_oracle.FlatOracle(counts, tables.alpha, tables.beta, threads=threads)
        splits, _, _ = fo.rounds(window_size, window_shift, constraint, num_rounds)
        sc = Scorer(counts, splits, tables)
        means, lmm = sc.mean_counts(), sc.log_marginal_likelyhoods()
        for k in range(len(splits) - 1):
            a, b = splits[k] + chrom_start, splits[k + 1] + chrom_start
            if output_mode == 'bedgraph':
                out.append('%s\t%d\t%d\t%f\n' % (chrom, a, b, means[k]))
            elif output_mode == 'bed':
                out.append('%s\t%d\t%d\n' % (chrom, a, b))
            elif output_mode == 'bedgraph+length+LMM':
                out.append('%s\t%d\t%d\t%f\t%d\t%f\n' % (chrom, a, b, means[k], splits[k + 1] - splits[k], lmm[k]))
            else:
                raise ValueError('Unknown output mode `%s`' % output_mode)
    return ''.join(out)
