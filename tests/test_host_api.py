"""Host-side logic that needs no GPU: the reference's fake-scorer tests
(/root/reference/tests/test_pasio.py:68-175, :230-325; tests/test_slice_when.py) restated against
pasio_b200, option validation, fusion-plan recognition, table builders, LPT sharding."""
import numpy as np
import pytest

from pasio_b200.splitters import (SquareSplitter, SlidingWindowReducer, RoundReducer, NotZeroReducer,
                                  NotConstantReducer, ReducerCombiner, NopSplitter, configure_splitter)
from pasio_b200.splitters import _fusion
from pasio_b200.dto.sliding_window import SlidingWindow
from pasio_b200.process_bedgraph import parse_bedgraph, parse_bedgraph_stream
from pasio_b200.utils.slice_when import slice_when
from pasio_b200.cached_log import LogComputer, LogGammaComputer
from pasio_b200 import sharding


class SimpleScorer:
    # user-defined scorer object (reference tests/test_pasio.py:68-87): not LogML, cannot run on a device
    def __init__(self, sequence, split_candidates):
        self.sequence = sequence
        self.split_candidates = split_candidates
        self.segment_creation_cost = 0

    def score(self, start, stop):
        return self.self_score(start, stop)

    def self_score(self, start, stop):
        start = self.split_candidates[start]
        stop = self.split_candidates[stop]
        if len(set(self.sequence[start:stop])) == 1:
            return (stop - start) ** 2
        return stop - start

    def all_suffixes_self_score(self, stop):
        return np.array([self.self_score(i, stop) for i in range(stop)], dtype='float64')


class SimpleGreedyScorer(SimpleScorer):
    def self_score(self, start, stop):
        return (self.split_candidates[stop] - self.split_candidates[start]) ** 0.5


simple = lambda counts, cands: SimpleScorer(counts, cands)
greedy = lambda counts, cands: SimpleGreedyScorer(counts, cands)


@pytest.mark.parametrize('seq,cands,splits,score', [
    ('A', None, [0, 1], 1), ('AAA', None, [0, 3], 9), ('AAABBB', None, [0, 3, 6], 18),
    ('AAABBBC', None, [0, 3, 6, 7], 19), ('ABBBC', None, [0, 1, 4, 5], 11),
    ('AAABBB', [0, 1, 2, 3, 5, 6], [0, 3, 6], 18), ('AAABBB', [0, 3, 5, 6], [0, 3, 6], 18),
    ('AAABBBC', [0, 3, 7], [0, 3, 7], 13), ('AAAAAA', [0, 3, 6], [0, 6], 36)])
def test_square_splitter_with_user_scorer(seq, cands, splits, score):
    cands = np.arange(len(seq) + 1) if cands is None else np.array(cands)
    got_score, got = SquareSplitter(simple).split(seq, cands)
    assert np.array_equal(got, splits) and got_score == score


def test_split_number_regularization():
    sp = SquareSplitter(SimpleScorer, split_number_regularization_multiplier=3,
                        split_number_regularization_function=lambda x: x)
    score, splits = sp.split('AAABAA', np.arange(7))
    assert np.array_equal(splits, [0, 3, 6]) and score == 9


def test_length_regularization():
    fn = lambda x: 1 / np.log(1 + x)
    sp = SquareSplitter(SimpleScorer, length_regularization_multiplier=1.5, length_regularization_function=fn)
    score, splits = sp.split('AAABAA', np.arange(7))
    assert np.array_equal(splits, [0, 3, 6])
    assert score == 9 + 3 - 1.5 * (1 / np.log(3 + 1) + 1 / np.log(3 + 1))
    score, splits = sp.split('AAABAA', np.array([0, 4, 5, 6]))
    assert np.array_equal(splits, [0, 4, 6])
    assert score == 4 + 4 - 1.5 * (1 / np.log(4 + 1) + 1 / np.log(2 + 1))


def test_collect_split_points():
    f = SquareSplitter.collect_split_points
    assert f([0, 0, 0]) == [0, 2]
    assert f([0, 0, 1, 2, 3, 4]) == [0, 1, 2, 3, 4, 5]
    assert f([0, 0, 0, 0, 0, 2, 3, 4]) == [0, 4, 7]
    assert f([0, 0, 0, 2, 1, 4]) == [0, 1, 4, 5]
    assert f([0, 0, 0, 2, 1, 3]) == [0, 2, 3, 5]


def test_sliding_window_with_user_scorer():
    A, B = 'A' * 16, 'B' * 17
    seq = A + B
    window = SlidingWindow(window_size=10, window_shift=5)
    base = SquareSplitter(simple)
    score, splits = ReducerCombiner(SlidingWindowReducer(window, base_reducer=base), base).split(seq, np.arange(len(seq) + 1))
    assert np.array_equal(splits, [0, len(A), len(seq)]) and score == len(A) ** 2 + len(B) ** 2
    base = SquareSplitter(simple, split_number_regularization_multiplier=2)
    score, splits = ReducerCombiner(SlidingWindowReducer(window, base_reducer=base), base).split(seq, np.arange(len(seq) + 1))
    assert np.array_equal(splits, [0, len(A), len(seq)]) and score == len(A) ** 2 + len(B) ** 2 - 2


def test_round_reducer_with_user_scorer():
    seq = 'A' * 40 + 'B' * 33 + 'C' * 50
    window = SlidingWindow(window_size=12, window_shift=6)
    base = SquareSplitter(simple)
    reducer = RoundReducer(SlidingWindowReducer(window, base))
    got = reducer.reduce_candidate_list(seq, np.arange(len(seq) + 1))
    once = SlidingWindowReducer(window, base).reduce_candidate_list(seq, got)
    assert np.array_equal(once, got)                       # fixed point
    assert 40 in got and 73 in got and got[0] == 0 and got[-1] == len(seq)
    one_round = RoundReducer(SlidingWindowReducer(window, base), num_rounds=1).reduce_candidate_list(seq, np.arange(len(seq) + 1))
    assert np.array_equal(one_round, SlidingWindowReducer(window, base).reduce_candidate_list(seq, np.arange(len(seq) + 1)))


def test_constant_reducers_on_plain_sequences():
    # inputs that are not coverage profiles (lists, float arrays) take the array-expression route on the host
    seq = [1, 1, 1, 2, 2, 2, 2]
    nc = NotConstantReducer()
    assert np.array_equal(nc.reduce_candidate_list(seq, np.array([0, 3, 7])), [0, 3, 7])
    assert np.array_equal(nc.reduce_candidate_list(seq, np.arange(8)), [0, 3, 7])
    assert np.array_equal(nc.reduce_candidate_list(np.array(seq, dtype=float), np.array([0, 3, 5, 7])), [0, 3, 7])
    assert np.array_equal(nc.reduce_candidate_list(seq, np.array([0, 5, 7])), [0, 7])
    assert np.array_equal(NotZeroReducer().reduce_candidate_list([0, 0, 0, 0, 0], np.arange(6)), [0, 5])
    assert np.array_equal(NotZeroReducer().reduce_candidate_list(seq, np.array([0, 5, 7])), [0, 5, 7])


def test_reducer_combiner_without_splitter():
    rc = ReducerCombiner(NotZeroReducer())
    with pytest.raises(Exception):
        rc.split(np.array([1]), np.array([0, 1]))
    with pytest.raises(Exception):
        rc.scorer(np.array([1]), np.array([0, 1]))


def test_sliding_window_geometry():
    # dto/sliding_window.py:9-15; trailing windows are subsets of earlier ones
    sizes = [len(w) for w, _ in SlidingWindow(100, 50).windows(np.arange(202))]
    assert sizes == [101, 101, 101, 52, 2]
    w = SlidingWindow(10, 5)
    assert w.ranges(2) == [(0, 2)]
    assert w.ranges(11) == [(0, 11), (5, 11)]
    assert [c for _, c in w.windows(np.arange(11))] == [1.0, 1.0]


def test_configure_splitter_validation_and_plans():
    with pytest.raises(ValueError):
        configure_splitter(algorithm='bogus')
    with pytest.raises(ValueError):
        configure_splitter(window_shift=None)
    with pytest.raises(ValueError):
        configure_splitter(window_size=None)
    with pytest.raises(ValueError):
        configure_splitter(length_regularization=1.0)
    with pytest.raises(ValueError):
        configure_splitter(length_regularization_function='revlog')
    with pytest.raises(ValueError):
        configure_splitter(split_constraints='bogus')
    s = configure_splitter(some_unknown_flag=1)
    plan = _fusion.pipeline_plan(s)
    assert plan['final'] == 'nop' and plan['steps'][0][0] == 'rounds' and plan['steps'][0][2:] == (2500, 1250, 'constants', None)
    assert _fusion.pipeline_plan(configure_splitter(algorithm='exact'))['final'] == 'exact'
    plan = _fusion.pipeline_plan(configure_splitter(algorithm='slidingwindow', split_constraints='zeros'))
    assert plan['final'] == 'exact' and plan['steps'][0][0] == 'window' and plan['steps'][0][4] == 'zeros'
    # regularised or user-scorer graphs are never fused
    assert _fusion.pipeline_plan(configure_splitter(split_number_regularization=1.0)) is None
    assert _fusion.pipeline_plan(SquareSplitter(simple)) is None
    s = configure_splitter(length_regularization=2.0, length_regularization_function='revlog', algorithm='exact')
    assert s.length_regularization_function(np.array([1.0])) == 1 / np.log(2.0)


def test_table_builders_extend_bit_identically():
    import scipy.special
    lc = LogComputer(shift=1.0, cache_size=4096)
    lg = LogGammaComputer(shift=2.5, cache_size=4096)
    n = 3 * (1 << 20) + 17            # crosses the chunked multi-thread build
    assert np.array_equal(lc.table(n)[:n], np.log(np.arange(n) + 1.0))
    assert np.array_equal(lg.table(n)[:n], scipy.special.gammaln(np.arange(n) + 2.5))
    x = np.array([0, 5, 4095, 4096, 100000, n - 1])
    assert np.array_equal(lc.compute_for_array_unbound(x), np.log(x + 1.0))
    assert np.array_equal(lg.compute_for_array(x, max_value=n), scipy.special.gammaln(x + 2.5))
    assert lc.compute_for_number(7) == np.log(8.0) and lc.compute_for_number(10 ** 7) == np.log(10 ** 7 + 1.0)


def test_approximate_log_gamma():
    # reference tests/test_pasio.py:178-190
    c = LogGammaComputer()
    tol = 1e-8
    for k in [256, 4095, 4096, 4097, 10000]:
        assert np.abs(np.log(np.arange(1, k + 1)).sum() - c.compute_for_number(k + 1)) < tol
    arr = np.array([0, 1, 20, 1024, 10000])
    want = np.array([np.log(np.arange(1, x + 1)).sum() for x in arr])
    assert np.allclose(c.compute_for_array_unbound(arr + 1), want, atol=tol)


def test_bedgraph_reader(tmp_path):
    # reference tests/test_pasio.py:237-260
    p = tmp_path / 'test.bedgraph'
    p.write_text('''chr1 0 10 0
        chr1 10 22 21
        chr1 22 23 30
        chr1 23 50 0
        chr2 0 15 0
        chr2 15 50 2
        chr2 50 60 0
        ''')
    chroms = {k: v for (k, v, l) in parse_bedgraph(str(p))}
    assert len(chroms) == 2 and len(chroms['chr1']) == 50 and len(chroms['chr2']) == 60
    assert np.all(chroms['chr1'][0:10] == 0) and np.all(chroms['chr1'][10:22] == 21)
    assert chroms['chr1'][22] == 30 and np.all(chroms['chr1'][23:50] == 0)
    assert np.all(chroms['chr2'][0:15] == 0) and np.all(chroms['chr2'][15:50] == 2) and np.all(chroms['chr2'][50:60] == 0)
    assert chroms['chr1'].dtype == int


def test_bedgraph_gaps(golden):
    import io
    text = 'c1\t5\t8\t2\nc1\t10\t12\t7.0\nc2\t0\t1\t4\n'
    filled = list(parse_bedgraph_stream(io.StringIO(text)))
    assert [(c, s) for c, _, s in filled] == [('c1', 5), ('c2', 0)]
    assert filled[0][1].tolist() == [2, 2, 2, 0, 0, 7, 7]
    cut = list(parse_bedgraph_stream(io.StringIO(text), split_at_gaps=True))
    assert [(c, p.tolist(), s) for c, p, s in cut] == [('c1', [2, 2, 2], 5), ('c1', [7, 7], 10), ('c2', [4], 0)]


def test_slice_when():
    # reference tests/test_slice_when.py
    neq = lambda a, b: a != b
    groups = lambda xs: [list(g) for g in slice_when(xs, neq)]
    assert groups([]) == []
    assert groups([1]) == [[1]]
    assert groups([1, 1, 2, 3, 3, 3]) == [[1, 1], [2], [3, 3, 3]]
    assert groups(iter([1, 2, 2])) == [[1], [2, 2]]
    it = slice_when([1, 1, 2, 2, 3], neq)
    first = next(it)
    assert next(first) == 1
    second = next(it)              # unfinished first group is drained
    assert list(second) == [2, 2]
    assert list(next(it)) == [3]
    with pytest.raises(StopIteration):
        next(it)


def test_lpt_assignment():
    lens = [248, 242, 198, 190, 181, 170, 159, 145, 138, 133, 135, 133, 114, 107, 101, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    for world in [1, 2, 4, 8]:
        rank_of = sharding.lpt_assign(lens, world)
        loads = [sum(l for l, r in zip(lens, rank_of) if r == k) for k in range(world)]
        assert sum(loads) == sum(lens)
        assert max(loads) <= sum(lens) / world + max(lens)        # LPT bound
        assert max(loads) / (sum(lens) / world) < 1.1 or world == 8
    # deterministic and order-independent of ties
    assert np.array_equal(sharding.lpt_assign(lens, 4), sharding.lpt_assign(list(lens), 4))
    res = sharding.segment_contigs([('a', np.zeros(3), 0), ('b', np.zeros(9), 0)], lambda n, c, s: n + str(len(c)))
    assert res == ['a3', 'b9']


def test_cpp_bedgraph_parser_equals_reference_semantics():
    """csrc/textio.cpp + process_bedgraph.contig_runs against the per-line restatement of the reference parser
    (oracle.parse_bedgraph_text): blank lines, mixed whitespace, float counts, gaps, overlaps, a chromosome that
    comes back later, a first interval ending at 0."""
    import io
    from oracle import pasio_oracle as po
    rs = np.random.RandomState(9)
    lines = ['', 'chr1\t5 8   2', 'chr1 8 12 7.0', '  ', 'chr1\t20\t25\t3e0', 'chr1 25 25 9', 'chr1 23 30 4', 'chr2 0 3 1',
             'chr1 30 33 2', 'chrZ -2 0 5', 'chrZ 0 4 6', 'chrZ 9 11 1\r']
    pos = 0
    for k in range(400):
        pos += int(rs.randint(0, 3)) * int(rs.randint(0, 50))
        ln = int(rs.randint(1, 40))
        lines.append('%s%s%d %d\t%d' % ('ctg%d' % (k // 97), ' \t'[k % 2], pos, pos + ln, int(rs.poisson(2))))
        pos += ln
    text = '\n'.join(lines) + '\n'
    for gaps in [False, True]:
        want = po.parse_bedgraph_text(text, split_at_gaps=gaps)
        got = list(parse_bedgraph_stream(io.StringIO(text), split_at_gaps=gaps))
        assert len(got) == len(want)
        for (c1, p1, s1), (c2, p2, s2) in zip(got, want):
            assert c1 == c2 and s1 == s2 and np.array_equal(p1, p2) and p1.dtype == int
    with pytest.raises(ValueError):
        list(parse_bedgraph_stream(io.StringIO('chr1 0 5\n')))
    with pytest.raises(ValueError):
        list(parse_bedgraph_stream(io.StringIO('chr1 a 5 1\n')))
    assert list(parse_bedgraph_stream(io.StringIO('\n\n'))) == []


def test_cpp_segment_formatter_equals_python_formatting():
    from pasio_b200 import _native
    rs = np.random.RandomState(1)
    splits = np.concatenate([[0], np.cumsum(rs.randint(1, 10 ** 6, 5000))]).astype(np.int64)
    means = np.concatenate([rs.gamma(1.0, 3.0, 4990), [0.0, 1e-7, 0.5000005, 123456789.1234565, 2.5e-7, 1e15, 0.1, 2.675, 1 / 3, 7.0]])
    lmm = -rs.gamma(2.0, 50.0, 5000)
    off = 12345
    want0 = ''.join('%s\t%d\t%d\t%f\n' % ('chrQ', a + off, b + off, m) for a, b, m in zip(splits[:-1], splits[1:], means))
    want1 = ''.join('%s\t%d\t%d\n' % ('chrQ', a + off, b + off) for a, b in zip(splits[:-1], splits[1:]))
    want2 = ''.join('%s\t%d\t%d\t%f\t%d\t%f\n' % ('chrQ', a + off, b + off, m, b - a, l)
                    for a, b, m, l in zip(splits[:-1], splits[1:], means, lmm))
    assert _native.format_segments('chrQ', off, splits, means, None, 0).decode() == want0
    assert _native.format_segments('chrQ', off, splits, None, None, 1).decode() == want1
    assert _native.format_segments('chrQ', off, splits, means, lmm, 2).decode() == want2


def test_threaded_text_io_equals_single_pass():
    """large inputs take the multi-threaded paths of csrc/textio.cpp (pieces cut at line / segment boundaries):
    same arrays as a per-line Python walk, same bytes as Python's % formatting; errors report the global line"""
    from pasio_b200 import _native
    rs = np.random.RandomState(4)
    # ~6 MB of text: blank lines, float counts, name changes right at and away from the piece seams
    n = 260000
    names = np.array(['chr%d' % (k // 17000) for k in range(n)])
    starts = np.cumsum(rs.randint(1, 30, n))
    stops = starts + rs.randint(1, 9, n)
    counts = rs.poisson(3, n)
    lines = []
    for k in range(n):
        c = '%d' % counts[k] if k % 1000 else '%d.0' % counts[k]
        lines.append('%s\t%d %d\t%s' % (names[k], starts[k], stops[k], c))
        if k % 7919 == 0:
            lines.append('   ')
    text = ('\n'.join(lines) + '\n').encode()
    assert len(text) > (4 << 20)
    r = _native.parse_bedgraph_text(text)
    assert np.array_equal(r['starts'], starts) and np.array_equal(r['stops'], stops) and np.array_equal(r['counts'], counts)
    got_names = np.array([text[o:o + l].decode() for o, l in zip(r['name_off'][::997], r['name_len'][::997])])
    assert np.array_equal(got_names, names[::997])
    want_new = np.concatenate([[1], (names[1:] != names[:-1]).astype(np.uint8)])
    assert np.array_equal(r['new_chrom'], want_new)
    assert r['n_float'] == len(range(0, n, 1000))
    # a malformed line late in the text: the error names its global line number (0-based, blank lines counted)
    bad_at = len(lines) - 5
    bad = list(lines)
    bad[bad_at] = 'chrX 1 2'
    with pytest.raises(ValueError) as ei:
        _native.parse_bedgraph_text(('\n'.join(bad) + '\n').encode())
    assert 'line %d ' % (bad_at + 1) in str(ei.value)

    m = 250000
    splits = np.concatenate([[0], np.cumsum(rs.randint(1, 5000, m))]).astype(np.int64)
    means = rs.gamma(1.0, 3.0, m)
    means[::501] = [0.5000005, 2.675, 1e-7][0]
    lmm = -rs.gamma(2.0, 50.0, m)
    want2 = ''.join('%s\t%d\t%d\t%f\t%d\t%f\n' % ('chr7', a + 9, b + 9, mu, b - a, l)
                    for a, b, mu, l in zip(splits[:-1].tolist(), splits[1:].tolist(), means.tolist(), lmm.tolist()))
    assert _native.format_segments('chr7', 9, splits, means, lmm, 2).decode() == want2
    want1 = ''.join('%s\t%d\t%d\n' % ('chr7', a + 9, b + 9) for a, b in zip(splits[:-1].tolist(), splits[1:].tolist()))
    assert _native.format_segments('chr7', 9, splits, None, None, 1).decode() == want1


def test_batch_segment_formatter_equals_per_contig_calls():
    """pasio_format_segments_batch (a batch of contigs segmented as one super-contig) writes the lines of per-contig
    pasio_format_segments calls, single- and multi-threaded (>= 100 000 segments)"""
    from pasio_b200 import _native
    rs = np.random.RandomState(8)
    for n_contigs, seg_hi in [(7, 40), (900, 400)]:
        nseg = rs.randint(1, seg_hi, n_contigs)
        offsets = [0]
        splits = [np.array([0], dtype=np.int64)]
        for c in range(n_contigs):
            inner = np.cumsum(rs.randint(1, 3000, nseg[c])).astype(np.int64)
            splits.append(offsets[-1] + inner)
            offsets.append(int(offsets[-1] + inner[-1]))
        splits = np.concatenate(splits)
        offsets = np.array(offsets, dtype=np.int64)
        first_split = np.searchsorted(splits, offsets)
        means = rs.gamma(1.0, 3.0, len(splits) - 1)
        lmm = -rs.gamma(2.0, 50.0, len(splits) - 1)
        chroms = ['ctg%d_%s' % (c, 'x' * (c % 5)) for c in range(n_contigs)]
        chrom_starts = rs.randint(0, 1000, n_contigs).astype(np.int64)
        for mode in (0, 1, 2):
            want = b''.join(
                _native.format_segments(chroms[c], int(chrom_starts[c]),
                                        splits[first_split[c]:first_split[c + 1] + 1] - offsets[c],
                                        means[first_split[c]:first_split[c + 1]] if mode != 1 else None,
                                        lmm[first_split[c]:first_split[c + 1]] if mode == 2 else None, mode)
                for c in range(n_contigs))
            got = _native.format_segments_batch(chroms, chrom_starts - offsets[:-1], first_split, splits,
                                                means if mode != 1 else None, lmm if mode == 2 else None, mode)
            assert got == want, (n_contigs, mode)
    assert len(splits) > 100000


def test_parse_bedgraph_from_files_plain_and_gzip(tmp_path):
    """files are opened in text mode like in the reference; the parser reads the bytes underneath"""
    import gzip
    import io
    from pasio_b200 import parse_bedgraph
    text = 'chr1\t0\t5\t2\nchr1\t5\t9\t0\nchr2 3 4 7\n\nchr2 10 12 1\n'
    want = [(c, p.tolist(), s) for c, p, s in parse_bedgraph_stream(io.StringIO(text))]
    assert want[0] == ('chr1', [2] * 5 + [0] * 4, 0) and want[1][0] == 'chr2'
    plain = tmp_path / 'a.bedgraph'
    plain.write_text(text)
    gz = tmp_path / 'a.bedgraph.gz'
    with gzip.open(str(gz), 'wt') as f:
        f.write(text)
    for fn in (plain, gz):
        assert [(c, p.tolist(), s) for c, p, s in parse_bedgraph(str(fn))] == want


def test_formatter_float_text_equals_python_percent_f():
    """'%f' through std::to_chars(fixed, 6): the characters Python's '%f' writes, for ties, tiny, huge, negative zero,
    random bit patterns"""
    from pasio_b200 import _native
    rs = np.random.RandomState(3)
    bits = rs.randint(0, 2 ** 63, 60000, dtype=np.int64).view(np.float64)
    vals = np.concatenate([bits[np.isfinite(bits)],
                           [0.0, -0.0, 0.5000005, 2.675, 1e-7, 2.5e-7, 5e-7, 1.5e-6, 2.5e-6, 0.9999995, 1e15, 1e22, 1e300,
                            -1e300, 5e-324, -4e-7, -5e-7, 123456789.1234565, float('inf'), float('-inf'), float('nan')],
                           rs.standard_normal(20000) * 10.0 ** rs.uniform(-8, 12, 20000)])
    splits = np.arange(len(vals) + 1, dtype=np.int64)
    got = _native.format_segments('c', 0, splits, vals, -vals, 2).decode().split('\n')[:-1]
    for k in range(len(vals)):
        assert got[k] == 'c\t%d\t%d\t%f\t%d\t%f' % (k, k + 1, vals[k], 1, -vals[k]), (k, vals[k])
