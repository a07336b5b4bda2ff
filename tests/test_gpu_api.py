"""The reference's own LogML tests (/root/reference/tests/test_pasio.py, test_bench_pasio.py) restated
against pasio_b200's drop-in classes, plus whole-pipeline parity with fixtures generated from the
unmodified reference.  Everything here computes on the GPU."""
import io
import math

import numpy as np
import pytest

import pasio_b200
from pasio_b200.log_marginal_likelyhood import (LogMarginalLikelyhoodIntAlphaComputer,
                                                LogMarginalLikelyhoodRealAlphaComputer, ScorerFactory)
from pasio_b200.splitters import (SquareSplitter, SlidingWindowReducer, RoundReducer, NotZeroReducer,
                                  NotConstantReducer, ReducerCombiner, NopSplitter, configure_splitter)
from pasio_b200.dto.sliding_window import SlidingWindow
from pasio_b200.process_bedgraph import split_bedgraph_stream
from pasio_b200 import synth
from oracle import pasio_oracle as po
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def test_stat_split_into_segments_square():
    # reference tests/test_pasio.py:12-43
    def split_on_two_segments_or_not(counts, scorer_factory):
        scorer = scorer_factory(counts, np.arange(len(counts) + 1))
        best_score = scorer.score(0, len(counts))
        split_point = 0
        for i in range(len(counts)):
            current_score = scorer.score(0, i) + scorer.score(i, len(counts))
            if current_score > best_score:
                split_point = i
                best_score = current_score
        return best_score, split_point

    np.random.seed(4)
    scorer_factory = lambda counts, split_candidates: LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, split_candidates)
    for repeat in range(5):
        counts = np.concatenate([np.random.poisson(15, 100), np.random.poisson(20, 100)])
        optimal_score, optimal_split = SquareSplitter(scorer_factory).split(counts, np.arange(len(counts) + 1))
        two_split_score, two_split_point = split_on_two_segments_or_not(counts, scorer_factory)
        assert optimal_score >= two_split_score
        assert two_split_point in optimal_split
        scorer = scorer_factory(counts, optimal_split)
        assert np.allclose(optimal_score, np.sum(scorer.scores()))
        assert abs(two_split_point - 100) < 10


def test_log_marginal_likelyhood_exact():
    # reference tests/test_pasio.py:45-65 (math.factorial instead of the removed np.math)
    def exact_function(counts, alpha, beta):
        fac = 1
        for c in counts:
            fac *= math.factorial(int(c))
        cs, ns = int(sum(counts)), len(counts)
        return np.log((beta ** alpha) * math.gamma(cs + alpha) / (math.gamma(alpha) * fac * ((ns + beta) ** (cs + alpha))))
    for counts, alpha, beta in [([0], 3, 5), ([0, 1], 3, 5), ([4, 0, 1, 3], 5, 2), ([4, 0, 1, 3], 1, 1)]:
        counts = np.array(counts)
        scorer = LogMarginalLikelyhoodIntAlphaComputer(counts, alpha, beta, split_candidates=np.array([0, len(counts)]))
        assert np.allclose(scorer.log_marginal_likelyhoods(), exact_function(counts, alpha, beta))


def test_suffixes_scores():
    # reference tests/test_pasio.py:193-204
    np.random.seed(2)
    counts = np.concatenate([np.random.poisson(15, 100), np.random.poisson(20, 100)])
    scorer = LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, np.arange(len(counts) + 1))
    assert np.allclose(scorer.all_suffixes_self_score(150), np.array([scorer.self_score(i, 150) for i in range(150)]))
    counts = np.array([0, 0, 1, 0, 0, 2, 2, 2, 10, 11, 100, 1, 0, 0, 1, 0], dtype='int64')
    scorer = LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, np.arange(len(counts) + 1))
    want = [scorer.self_score(i, len(counts) - 1) for i in range(len(counts) - 1)]
    assert np.allclose(scorer.all_suffixes_self_score(len(counts) - 1), np.array(want))


def test_suffixes_scores_with_candidates():
    # reference tests/test_pasio.py:206-224
    np.random.seed(2)
    counts = np.arange(1, 10)
    scorer = LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, np.arange(len(counts) + 1))
    candidates = np.array([0, 1, 3, 4, 5, 6, 7, 8, 9])
    scorer_with_candidates = LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, candidates)
    assert np.allclose(scorer.all_suffixes_self_score(9)[candidates[:-1]],
                       scorer_with_candidates.all_suffixes_self_score(8))
    counts = np.concatenate([np.random.poisson(15, 100), np.random.poisson(20, 100)])
    scorer = LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, np.arange(len(counts) + 1))
    candidates = np.array([0, 1, 10, 20, 21, 30, 40, 149, 200])
    scorer_with_candidates = LogMarginalLikelyhoodIntAlphaComputer(counts, 1, 1, candidates)
    assert np.allclose(scorer.all_suffixes_self_score(200)[candidates[:-1]],
                       scorer_with_candidates.all_suffixes_self_score(len(candidates) - 1))


def test_scorer_attributes_match_oracle():
    counts = synth.piecewise_poisson(3000, 8)
    cands = synth.random_candidates(3000, 200, 8)
    for alpha, beta, cls in [(1, 1, LogMarginalLikelyhoodIntAlphaComputer), (2.5, 0.5, LogMarginalLikelyhoodRealAlphaComputer)]:
        sc = cls(counts, alpha, beta, cands)
        ref = po.Scorer(counts, cands, po.Tables(alpha, beta))
        assert np.array_equal(sc.cumsum, ref.cumsum)
        assert sc.segment_creation_cost == ref.pen
        assert np.array_equal(sc.scores(), ref.scores())
        assert np.array_equal(sc.mean_counts(), ref.mean_counts())
        # logfac_cumsum is summed sequentially like np.cumsum (csrc/logfac_exact.cu): bit for bit
        assert np.array_equal(sc.logfac_cumsum, ref.logfac_cumsum)
        assert np.array_equal(sc.log_marginal_likelyhoods(), ref.log_marginal_likelyhoods())
        assert sc.total_sum_logfac() == ref.total_sum_logfac()
        assert sc.score(3, 17) == ref.row(17)[3] + ref.pen
        assert sc.score_no_splits() == ref.row(len(cands) - 1)[0] + ref.pen


def test_scorer_asserts():
    with pytest.raises(AssertionError):
        LogMarginalLikelyhoodIntAlphaComputer(np.array([1, -2, 3]), 1, 1, np.array([0, 3]))
    with pytest.raises(AssertionError):
        LogMarginalLikelyhoodIntAlphaComputer(np.array([1.0, 2.0]), 1, 1, np.array([0, 2]))
    with pytest.raises(AssertionError):
        LogMarginalLikelyhoodIntAlphaComputer(np.array([1, 2, 3]), 1, 1, np.array([0, 2]))
    with pytest.raises(AssertionError):
        LogMarginalLikelyhoodIntAlphaComputer(np.array([1, 2, 3]), 1, 1, np.array([0, 2, 2, 3]))
    with pytest.raises(AssertionError):
        LogMarginalLikelyhoodIntAlphaComputer([1, 2, 3], 1, 1, np.array([0, 3]))
    with pytest.raises(AssertionError):
        ScorerFactory(-1, 1)


def test_benchmark_shapes():
    # reference tests/test_bench_pasio.py:19-48 shapes (results vs the oracle)
    f = lambda c, s: LogMarginalLikelyhoodIntAlphaComputer(c, 1, 1, s)
    t = po.Tables(1, 1)
    for half in [50, 500]:
        np.random.seed(2)
        counts = np.concatenate([np.random.poisson(15, half), np.random.poisson(20, half)])
        score, splits = SquareSplitter(f).split(counts, np.arange(len(counts) + 1))
        o = po.square_split(counts, np.arange(len(counts) + 1), t)
        assert score == o[0] and np.array_equal(splits, o[1])
    np.random.seed(2)
    counts = np.concatenate([np.random.poisson(15, 50000), np.random.poisson(20, 50000)])
    cands = np.hstack([np.arange(0, len(counts), 100), 100000])
    score, splits = SquareSplitter(f).split(counts, cands)
    o = po.square_split(counts, cands, t)
    assert score == o[0] and np.array_equal(splits, o[1])


def test_object_route_equals_fused_route():
    """lambda factories are not fused: windows go one by one through the objects (and still run the DP
    on the GPU); the result must equal the fused one-launch-per-round path."""
    counts = synth.dnase_like(12000, 21, hotspot_share=0.5)
    cands = np.arange(len(counts) + 1)
    lam = lambda c, s: LogMarginalLikelyhoodIntAlphaComputer(c, 1, 1, s)
    fused_f = ScorerFactory(1.0, 1.0)
    for head in [NotConstantReducer, NotZeroReducer, None]:
        def graph(fac):
            sq = SquareSplitter(fac)
            base = ReducerCombiner(head(), sq) if head else sq
            return RoundReducer(SlidingWindowReducer(SlidingWindow(200, 100), base))
        slow = graph(lam).reduce_candidate_list(counts, cands)
        fast = graph(fused_f).reduce_candidate_list(counts, cands)
        assert np.array_equal(slow, fast)
        one = SlidingWindowReducer(SlidingWindow(200, 100), SquareSplitter(fused_f)).reduce_candidate_list(counts, cands)
        assert one[0] == 0 and one[-1] == len(counts)


@pytest.mark.parametrize('name,kwargs', [
    ('default300k', dict()),
    ('exact4k', dict(algorithm='exact')),
    ('zeros100k', dict(split_constraints='zeros', window_size=600, window_shift=300)),
    ('real100k', dict(alpha=0.7, beta=2.0, window_size=1000, window_shift=500)),
    ('rounds2', dict(num_rounds=2, window_size=1000, window_shift=500)),
])
def test_segments_with_scores_vs_reference_fixture(golden, name, kwargs):
    g = golden('pipeline.npz')
    counts = g[name + '.counts'].astype(np.int64)
    splitter = configure_splitter(**kwargs)
    segs = list(pasio_b200.segments_with_scores(counts, splitter))
    starts = np.array([s.start for s in segs])
    stops = np.array([s.stop for s in segs])
    means = np.array([s.mean_count for s in segs])
    lmm = np.array([s.log_marginal_likelyhood for s in segs])
    score, splits = splitter.split(counts, np.arange(len(counts) + 1))
    g.check_splits(np.concatenate([starts, stops[-1:]]), np.concatenate([g[name + '.starts'], g[name + '.stops'][-1:]]),
                   score, g[name + '.score'], name)
    assert np.array_equal(splits[:-1], starts)
    if g.bit_exact(name):
        assert np.array_equal(means, g[name + '.mean'])
        # logfac_cumsum is summed in the reference's order (csrc/logfac_exact.cu): the LMM column is bit-identical
        assert np.array_equal(lmm, g[name + '.lmm'])


def test_split_bedgraph_text_vs_reference_fixture(golden):
    g = golden('pipeline.npz')
    text = str(g['bg.input'])
    for mode in ['bedgraph', 'bed', 'bedgraph+length+LMM']:
        for gaps in [False, True]:
            out = io.StringIO()
            split_bedgraph_stream(io.StringIO(text), out, configure_splitter(window_size=500, window_shift=250),
                                  split_at_gaps=gaps, output_mode=mode)
            want = str(g['bg.%s.%d' % (mode, int(gaps))])
            if not g.bit_exact('bedgraph text %s gaps=%d' % (mode, gaps)):
                pytest.skip('host np.log/gammaln tables differ from the fixtures\' (see the pytest header): the text '
                            'comparison with the reference fixture cannot be made on this host')
            assert out.getvalue() == want, (mode, gaps)          # all three modes byte for byte, the LMM column included


def test_slidingwindow_algorithm_graph():
    # the reference crashes for algorithm='slidingwindow' (NameError); the intended graph is
    # ReducerCombiner(SlidingWindowReducer, SquareSplitter) -- checked against the oracle
    counts = synth.dnase_like(20000, 3, hotspot_share=0.5)
    splitter = configure_splitter(algorithm='slidingwindow', window_size=300, window_shift=150)
    score, splits = splitter.split(counts, np.arange(len(counts) + 1))
    t = po.Tables(1, 1.0)
    reduced = po.sliding_window_round(counts, np.arange(len(counts) + 1), t, 300, 150, 'constants')
    o = po.square_split(counts, reduced, t)
    assert score == o[0] and np.array_equal(splits, o[1])
    segs = list(pasio_b200.segments_with_scores(counts, splitter))
    assert np.array_equal([s.start for s in segs], o[1][:-1])


def test_regularized_split_uses_device_rows():
    """regularisation callables are Python: the recurrence is host-driven, rows come from the GPU"""
    counts = synth.piecewise_poisson(400, 4)
    cands = np.arange(401)
    sp = SquareSplitter(ScorerFactory(1.0, 1.0), split_number_regularization_multiplier=3.0)
    score, splits = sp.split(counts, cands)
    o = po.square_split_regularized(counts, cands, po.Tables(1, 1.0), num_mult=3.0)
    assert score == o[0] and np.array_equal(splits, o[1])
    sp = SquareSplitter(ScorerFactory(1.0, 1.0), length_regularization_multiplier=1.5,
                        length_regularization_function=lambda x: 1 / np.log(1 + x))
    score, splits = sp.split(counts, cands)
    o = po.square_split_regularized(counts, cands, po.Tables(1, 1.0), len_mult=1.5, len_fn=lambda x: 1 / np.log(1 + x))
    assert score == o[0] and np.array_equal(splits, o[1])


def test_regularized_dp_on_device_vs_oracle(monkeypatch):
    """--split-number-regularization / --length-regularization run as ONE kernel per candidate list
    (csrc/regularized_dp.cu), no per-row launches, bit-equal to the reference recurrence at N = 20 001"""
    from pasio_b200 import _native
    from pasio_b200.splitters.square_splitter import _revlog

    def no_rows(self, stop):
        raise AssertionError('the regularised DP fetched a row from the device: the host loop is still in use')
    monkeypatch.setattr(_native.Engine, 'suffix_scores', no_rows)
    counts = synth.piecewise_poisson(20000, 4)
    allpos = np.arange(len(counts) + 1)
    sparse = synth.random_candidates(len(counts), 3000, 8)
    revlog = lambda x: 1 / np.log(x + 1)                                        # noqa: E731
    cases = [
        (allpos, (1.0, 1.0), dict(split_number_regularization_multiplier=3.0), dict(num_mult=3.0)),
        (allpos, (1.0, 1.0), dict(length_regularization_multiplier=1.5, length_regularization_function=_revlog),
         dict(len_mult=1.5, len_fn=revlog)),
        (sparse, (2.5, 3.0), dict(length_regularization_multiplier=40.0, length_regularization_function=_revlog,
                                  split_number_regularization_multiplier=0.75), dict(len_mult=40.0, len_fn=revlog, num_mult=0.75)),
        (sparse, (1.0, 1.0), dict(length_regularization_multiplier=0.001), dict(len_mult=0.001)),      # identity length penalty
        (allpos[:302], (1.0, 1.0), dict(split_number_regularization_multiplier=2), dict(num_mult=2)),  # integer multiplier
    ]
    for cands, ab, kw, okw in cases:
        c = counts[:cands[-1]]
        score, splits = SquareSplitter(ScorerFactory(*ab), **kw).split(c, cands)
        o = po.square_split_regularized(c, cands, po.Tables(po.normalise_alpha(ab[0]), ab[1]), **okw)
        assert score == o[0] and np.array_equal(splits, o[1]), kw
    # through configure_splitter, as the CLI builds it (the regularised splitter is also the windows' base reducer)
    splitter = configure_splitter(algorithm='exact', length_regularization=1.5, length_regularization_function='revlog',
                                  split_number_regularization=0.5)
    score, splits = splitter.split(counts[:5000], np.arange(5001))
    o = po.square_split_regularized(counts[:5000], np.arange(5001), po.Tables(1, 1.0), len_mult=1.5, len_fn=revlog, num_mult=0.5)
    assert score == o[0] and np.array_equal(splits, o[1])
    # an arbitrary callable need not be element-wise: host route (rows from the device)
    monkeypatch.undo()
    sp = SquareSplitter(ScorerFactory(1.0, 1.0), length_regularization_multiplier=1.5, length_regularization_function=revlog)
    score, splits = sp.split(counts[:400], np.arange(401))
    o = po.square_split_regularized(counts[:400], np.arange(401), po.Tables(1, 1.0), len_mult=1.5, len_fn=revlog)
    assert score == o[0] and np.array_equal(splits, o[1])


# ---- the reference's NotZero / NotConstant tests (tests/test_pasio.py:304-325): int profiles run on the GPU ----
class _SimpleScorer:
    def __init__(self, sequence, split_candidates):
        self.sequence = sequence
        self.split_candidates = split_candidates
        self.segment_creation_cost = 0

    def self_score(self, start, stop):
        start = self.split_candidates[start]
        stop = self.split_candidates[stop]
        if len(set(self.sequence[start:stop])) == 1:
            return (stop - start) ** 2
        return stop - start

    def all_suffixes_self_score(self, stop):
        return np.array([self.self_score(i, stop) for i in range(stop)], dtype='float64')


class _GreedyScorer(_SimpleScorer):
    def self_score(self, start, stop):
        return (self.split_candidates[stop] - self.split_candidates[start]) ** 0.5


_simple = lambda counts, cands: _SimpleScorer(counts, cands)
_greedy = lambda counts, cands: _GreedyScorer(counts, cands)


def test_not_constant_and_not_zero():
    seq = np.array([1, 1, 1, 2, 2, 2, 2])
    allp = np.arange(len(seq) + 1)
    assert np.array_equal(ReducerCombiner(NotZeroReducer(), SquareSplitter(_simple)).split(seq, allp)[1], [0, 3, 7])
    assert np.array_equal(ReducerCombiner(NotZeroReducer(), SquareSplitter(_greedy)).split(seq, allp)[1], list(range(8)))
    assert np.array_equal(ReducerCombiner(NotConstantReducer(), SquareSplitter(_greedy)).split(seq, allp)[1], [0, 3, 7])
    assert np.array_equal(ReducerCombiner(NotConstantReducer(), SquareSplitter(_greedy)).split(seq, np.array([0, 1, 2, 3, 4, 5, 7]))[1], [0, 3, 7])
    nc = NotConstantReducer()
    assert np.array_equal(nc.reduce_candidate_list(seq, np.array([0, 3, 7])), [0, 3, 7])
    assert np.array_equal(nc.reduce_candidate_list(seq, np.arange(8)), [0, 3, 7])
    assert np.array_equal(nc.reduce_candidate_list(seq, np.array([0, 3, 5, 7])), [0, 3, 7])
    assert np.array_equal(nc.reduce_candidate_list(seq, np.array([0, 5, 7])), [0, 7])
    assert np.array_equal(NotZeroReducer().reduce_candidate_list(np.zeros(5, dtype=int), np.arange(6)), [0, 5])
    assert np.array_equal(NotZeroReducer().reduce_candidate_list(seq, np.array([0, 5, 7])), [0, 5, 7])


def test_reducers_vs_reference_fixture(golden):
    g = golden('reducers.npz')
    for k in range(6):
        c, cands = g['r%d.counts' % k], g['r%d.cands' % k]
        assert np.array_equal(NotZeroReducer().reduce_candidate_list(c, cands), g['r%d.notzero' % k])
        assert np.array_equal(NotConstantReducer().reduce_candidate_list(c, cands), g['r%d.notconstant' % k])



def test_constant_reducers_random_vs_oracle():
    rs = np.random.RandomState(3)
    for _ in range(10):
        n = int(rs.randint(2, 5000))
        counts = (rs.poisson(0.5, n) * (rs.random_sample(n) < 0.5)).astype(np.int64)
        inner = np.sort(rs.choice(np.arange(1, n), size=int(rs.randint(0, n - 1)), replace=False)) if n > 2 else np.array([], dtype=int)
        cands = np.concatenate([[0], inner, [n]]).astype(np.int64)
        assert np.array_equal(NotConstantReducer().reduce_candidate_list(counts, cands), po.not_constant(counts, cands))
        assert np.array_equal(NotZeroReducer().reduce_candidate_list(counts, cands), po.not_zero(counts, cands))
    z = np.zeros(100, dtype=np.int64)
    assert np.array_equal(NotZeroReducer().reduce_candidate_list(z, np.arange(101)), [0, 100])
    assert np.array_equal(NotConstantReducer().reduce_candidate_list(z, np.arange(101)), [0, 100])


def test_window_larger_than_one_cta_goes_through_exact_kernel():
    """window_size beyond the fused limit is not an error at the API level: windows run one by one on the exact-DP kernel"""
    from pasio_b200.splitters import _fusion
    counts = synth.piecewise_poisson(12000, 5) + 1           # every position is a change point
    splitter = configure_splitter(window_size=9000, window_shift=4500, split_constraints='none', num_rounds=1)
    assert _fusion.pipeline_plan(splitter) is None
    score, splits = splitter.split(counts, np.arange(len(counts) + 1))
    t = po.Tables(1, 1.0)
    want = po.sliding_window_round(counts, np.arange(len(counts) + 1), t, 9000, 4500, 'none')
    assert np.array_equal(splits, want)


def test_deep_coverage_total_beyond_int32():
    """prefix sums are int64: a contig whose total count exceeds 2^31 runs through the default pipeline
    (the 32-bit limit applies to the count inside one window's DP, not to the contig)"""
    rs = np.random.RandomState(6)
    counts = np.repeat(rs.poisson(110, 400000), 50).astype(np.int64)          # 20 Mb, runs of 50 nt
    assert counts.sum() > 2 ** 31
    splitter = configure_splitter(window_size=500, window_shift=250)
    score, splits = splitter.split(counts, np.arange(len(counts) + 1))
    want, _, _ = c_oracle.FlatOracle(counts, 1.0, 1.0).rounds(500, 250, 'constants')
    assert np.array_equal(splits, want)


def test_split_bedgraph_batches_short_contigs():
    """many short contigs go to the device as one batch per launch: byte-identical to contig-by-contig processing,
    in input order, for the three output modes; a long contig in between runs alone; exact mode is never batched"""
    import io
    from pasio_b200 import process_bedgraph as pb
    rs = np.random.RandomState(12)
    lines = []
    for c in range(240):
        n = int(rs.randint(1000, 30000)) if c != 100 else 400000
        counts = synth.dnase_like(n, 700 + c, hotspot_share=0.3)
        lines.extend(synth.to_bedgraph_lines('tx%d' % c, counts, chrom_start=int(rs.randint(0, 50))))
    text = ''.join(lines)

    def run(splitter, mode, batch_nt, alone=pb.BATCH_ALONE):
        old = pb.BATCH_NT, pb.BATCH_ALONE
        pb.BATCH_NT, pb.BATCH_ALONE = batch_nt, alone
        try:
            out = io.StringIO()
            split_bedgraph_stream(io.StringIO(text), out, splitter, output_mode=mode)
            return out.getvalue()
        finally:
            pb.BATCH_NT, pb.BATCH_ALONE = old

    splitter = configure_splitter(window_size=800, window_shift=400)
    for mode in ['bedgraph', 'bed', 'bedgraph+length+LMM']:
        one_by_one = run(splitter, mode, 0)
        assert run(splitter, mode, 1 << 27, alone=300000) == one_by_one          # contig 100 alone, the others batched
        assert run(splitter, mode, 200000) == one_by_one                        # several small batches
    assert one_by_one.count('\n') > 240
    exact = configure_splitter(algorithm='exact')
    short_text = ''.join(l for l in lines if l.split('\t')[0] in ('tx3', 'tx4'))
    out = io.StringIO()
    split_bedgraph_stream(io.StringIO(short_text), out, exact)
    assert out.getvalue().startswith('tx3\t')


def _subsampled_text(kind, n_contigs, seed):
    """a sub-sample of BASELINE config 4 (hg38 contig-size profile, scaled) / config 5 (transcript-like stream) as
    run-length bedgraph text"""
    rs = np.random.RandomState(seed)
    lines = []
    if kind == 'genome':
        sizes = synth.genome_profile(scale=0.004)                    # 195 contigs, 1 kb .. 1 Mb
        pick = sorted(rs.choice(len(sizes), size=n_contigs, replace=False))
        for i in pick:
            name, n = sizes[i]
            lines.extend(synth.to_bedgraph_lines(name, synth.dnase_like(n, seed=i), chrom_start=int(rs.randint(0, 1000))))
    else:
        lens = synth.transcript_lengths(n_contigs, seed=seed)
        for c, n in enumerate(lens):
            lines.extend(synth.to_bedgraph_lines('tx%05d' % c, synth.dnase_like(int(n), seed=5000 + c, hotspot_share=0.3)))
    return ''.join(lines)


@pytest.mark.parametrize('kind,n_contigs', [('genome', 120), ('transcripts', 1000)])
def test_configs_4_and_5_subsample_text_vs_oracle(kind, n_contigs):
    """BASELINE configs 4 / 5 at sub-sampled shape through the text boundary (split_bedgraph_stream: chunked reader,
    batched launches, C++ formatter) against the oracle's parser + pipeline + '%' formatting, byte for byte."""
    import io
    from pasio_b200 import process_bedgraph as pb
    text = _subsampled_text(kind, n_contigs, 77)
    want = po.split_bedgraph_text(text, po.Tables(1, 1.0), threads=8)
    old = pb.CHUNK_BYTES
    pb.CHUNK_BYTES = 1 << 20                                          # many pieces: contigs straddle the seams
    try:
        out = io.StringIO()
        split_bedgraph_stream(io.StringIO(text), out, configure_splitter())
        got = out.getvalue()
    finally:
        pb.CHUNK_BYTES = old
    assert got.count('\n') == want.count('\n')
    assert got == want
    for mode in ['bed', 'bedgraph+length+LMM']:        # the LMM column: every contig of a batch restarts its log-factorial sum
        out = io.StringIO()
        split_bedgraph_stream(io.StringIO(text), out, configure_splitter(), output_mode=mode)
        assert out.getvalue() == po.split_bedgraph_text(text, po.Tables(1, 1.0), output_mode=mode, threads=8), mode


def test_split_bedgraph_devices_pool_equals_single_process():
    """devices=N: one worker process per GPU, LPT over the batches, shard files gathered in input order -- the same text
    (with one GPU in the box both workers share it; the partition, the spawn path and the gather are the same)"""
    import io
    import ctypes
    from pasio_b200 import _native, device_pool
    text = _subsampled_text('genome', 40, 5)
    single = io.StringIO()
    split_bedgraph_stream(io.StringIO(text), single, configure_splitter(window_size=800, window_shift=400),
                          output_mode='bedgraph+length+LMM')
    import torch
    n_dev = max(1, min(2, torch.cuda.device_count()))
    old_worker = device_pool._worker
    multi = io.StringIO()
    if n_dev == 1:                                                    # two workers on the one GPU
        import os
        os.environ['PASIO_B200_POOL_SAME_DEVICE'] = '1'
    try:
        split_bedgraph_stream(io.StringIO(text), multi, configure_splitter(window_size=800, window_shift=400),
                              output_mode='bedgraph+length+LMM', devices=2)
    finally:
        import os
        os.environ.pop('PASIO_B200_POOL_SAME_DEVICE', None)
    assert multi.getvalue() == single.getvalue()
    assert single.getvalue().count('\n') > 40


def test_contig_load_device_equals_host_load():
    """pasio_contig_load_device (counts already in HBM: the entry point bench.py's `value` runs through) == pasio_contig_load
    of the same counts: prefix sums, rounds, scores; single contig and a batch with offsets; alignment is checked"""
    import torch
    from pasio_b200 import _native
    from pasio_b200.log_marginal_likelyhood import ScorerFactory
    eng = _native.engine()
    f = ScorerFactory(1.0, 1.0)
    eng.use_scorer(f)
    counts = synth.dnase_like(3000000, 21, hotspot_share=0.2)
    dev = torch.from_numpy(counts).cuda()
    for offsets in [None, np.array([0, 1000000, 1000001, 2500000, 3000000], dtype=np.int64)]:
        eng.invalidate()
        eng.load(counts, offsets=offsets)
        eng.set_candidates(None)
        sizes_a, final_a, cells_a = eng.rounds(2500, 1250, 'constants')
        splits_a = eng.candidates().copy()
        scores_a = eng.segment_scores(scores=True)[0].copy()
        cum_a = eng.cumsum_at_candidates()
        eng.invalidate()
        eng.load_device(dev.data_ptr(), len(counts), owner=dev, offsets=offsets)
        assert eng.info()[0] == len(counts) and eng.info()[1] == int(counts.sum())
        eng.set_candidates(None)
        sizes_b, final_b, cells_b = eng.rounds(2500, 1250, 'constants')
        assert (sizes_b, final_b, cells_b) == (sizes_a, final_a, cells_a)
        assert np.array_equal(eng.candidates(), splits_a)
        assert np.array_equal(eng.segment_scores(scores=True)[0], scores_a)
        assert np.array_equal(eng.cumsum_at_candidates(), cum_a)
    want, _, _ = c_oracle.FlatOracle(counts, 1.0, 1.0, threads=8).rounds(2500, 1250, 'constants')
    eng.invalidate()
    eng.load_device(dev.data_ptr(), len(counts), owner=dev)
    eng.set_candidates(None)
    eng.rounds(2500, 1250, 'constants')
    assert np.array_equal(eng.candidates(), want)
    with pytest.raises(ValueError):
        eng.load_device(dev.data_ptr() + 8, len(counts) - 1, owner=dev)          # not 16-byte aligned
    neg = dev.clone()
    neg[12345] = -5
    with pytest.raises(AssertionError):
        eng.load_device(neg.data_ptr(), len(counts), owner=neg)


def test_window_prune_switch_gives_identical_rounds():
    """the far-column bound of the window kernel and the skipping of covered windows are exact: with both switched off
    (every cell of every window evaluated) every round of a bench-like contig gives the same candidates"""
    from pasio_b200 import _native
    from pasio_b200.log_marginal_likelyhood import ScorerFactory
    eng = _native.engine()
    eng.use_scorer(ScorerFactory(1.0, 1.0))
    counts = synth.dnase_like(30000000, 1000)                        # the first 30 Mb of the bench contig
    eng.invalidate()
    eng.load(counts)
    per_round = {}
    try:
        for prune in (1, 0):
            eng.set_tuning('window_prune', prune)
            eng.set_tuning('window_phases', prune)
            eng.set_candidates(None)
            rounds = []
            while True:
                n_in, n_out, cells = eng.round(2500, 1250, 'constants')
                rounds.append((n_in, n_out, cells, eng.round_stats()[1], eng.candidates().copy()))
                if n_in == n_out:
                    break
            per_round[prune] = rounds
    finally:
        eng.set_tuning('window_prune', 1)
        eng.set_tuning('window_phases', 1)
    assert len(per_round[0]) == len(per_round[1])
    for a, b in zip(per_round[1], per_round[0]):
        assert a[:3] == b[:3] and np.array_equal(a[4], b[4])
        assert b[3] == 0                                              # nothing skipped with the switches off
    assert sum(a[3] for a in per_round[1]) > 0


def test_window_speculation_switch_gives_identical_rounds():
    """the speculative block resolution of the window kernel (assume arg-max = previous row, verify, else the ordinary
    chain) writes the same P and prev: every round of a bench-like contig gives the same candidates with it on and off,
    and the oracle agrees with the final list"""
    from pasio_b200 import _native
    from pasio_b200.log_marginal_likelyhood import ScorerFactory
    eng = _native.engine()
    eng.use_scorer(ScorerFactory(1.0, 1.0))
    counts = synth.dnase_like(20000000, 1000)
    eng.invalidate()
    eng.load(counts)
    per_round = {}
    try:
        for spec in (1, 0):
            eng.set_tuning('window_speculate', spec)
            eng.set_candidates(None)
            rounds = []
            while True:
                n_in, n_out, cells = eng.round(2500, 1250, 'constants')
                rounds.append((n_in, n_out, cells, eng.candidates().copy()))
                if n_in == n_out:
                    break
            per_round[spec] = rounds
    finally:
        eng.set_tuning('window_speculate', 1)
    assert len(per_round[0]) == len(per_round[1]) and len(per_round[1]) >= 4
    for a, b in zip(per_round[1], per_round[0]):
        assert a[:3] == b[:3] and np.array_equal(a[3], b[3])
    # a real alpha takes the other template instance
    eng.use_scorer(ScorerFactory(0.75, 1.5))
    small = synth.dnase_like(3000000, 7)
    eng.invalidate()
    eng.load(small)
    finals = {}
    try:
        for spec in (1, 0):
            eng.set_tuning('window_speculate', spec)
            eng.set_candidates(None)
            eng.rounds(2500, 1250, 'constants')
            finals[spec] = eng.candidates().copy()
    finally:
        eng.set_tuning('window_speculate', 1)
    assert np.array_equal(finals[0], finals[1])
    want, _, _ = c_oracle.FlatOracle(small, 0.75, 1.5, threads=8).rounds(2500, 1250, 'constants')
    assert np.array_equal(finals[1], want)


def test_logfac_exact_and_parallel_modes():
    """logfac_cumsum: the default reproduces np.cumsum's sequential float64 sum bit for bit (sparse and dense coverage, a
    batch of contigs: every contig's sum restarts at 0); PASIO_TUNE_LOGFAC_EXACT = 0 (parallel scan) agrees to 1e-9"""
    from pasio_b200 import _native
    eng = _native.engine()
    f = ScorerFactory(1.0, 1.0)
    t = po.Tables(1, 1.0)
    for counts in [synth.dnase_like(400000, 3, hotspot_share=0.3), synth.two_level_poisson(60000, seed=5),
                   np.zeros(5000, dtype=np.int64), np.ones(777, dtype=np.int64)]:
        cands = synth.random_candidates(len(counts), min(3000, len(counts) // 2), 9)
        ref = po.Scorer(counts, cands, t)
        eng.use_scorer(f)
        eng.invalidate()
        eng.load(counts)
        eng.set_candidates(cands)
        try:
            for exact in (1, 0):
                eng.set_tuning('logfac_exact', exact)
                lmm, total = eng.segment_lmm()
                lf = eng.segment_scores(scores=False, logfac=True)[3]
                if exact:
                    assert np.array_equal(lf, ref.logfac_cumsum) and np.array_equal(lmm, ref.log_marginal_likelyhoods())
                    assert total == ref.total_sum_logfac()
                else:
                    assert np.allclose(lf, ref.logfac_cumsum, rtol=1e-12, atol=1e-9)
                    assert np.allclose(lmm, ref.log_marginal_likelyhoods(), rtol=1e-9, atol=1e-7)
        finally:
            eng.set_tuning('logfac_exact', 1)
    # a batch: three contigs as one super-contig with boundaries
    parts = [synth.dnase_like(50000, 11, hotspot_share=0.4), synth.two_level_poisson(4000, seed=6), synth.dnase_like(20000, 12)]
    batch = np.concatenate(parts)
    offsets = np.concatenate([[0], np.cumsum([len(x) for x in parts])]).astype(np.int64)
    eng.invalidate()
    eng.load(batch, offsets=offsets)
    eng.set_candidates(None)
    eng.rounds(500, 250, 'constants')
    splits = eng.candidates()
    lmm, _ = eng.segment_lmm()
    want = []
    for k, part in enumerate(parts):
        local = splits[(splits >= offsets[k]) & (splits <= offsets[k + 1])] - offsets[k]
        want.append(po.Scorer(part, local, t).log_marginal_likelyhoods())
    assert np.array_equal(lmm, np.concatenate(want))


def test_logfac_sums_behind_the_upload():
    """load_and_round(want_logfac=True) forms the sequential log-factorial sums chunk by chunk behind the upload (40 Mb = two
    chunks): logfac_cumsum, the LMM column and the total are the numpy values bit for bit, the same as on demand; a count
    beyond the lgamma table makes the library discard those sums and take the ordinary path (which grows the table)"""
    from pasio_b200 import _native
    eng = _native.engine()
    t = po.Tables(1, 1.0)
    for n, poke in [(40000000, None), (3000000, None), (700000, 1500000)]:
        counts = synth.dnase_like(n, 21, hotspot_share=0.3)
        if poke:
            counts[n // 2] = poke                       # gammaln table of a fresh factory has 2^20 entries
        results = []
        for eager in (True, False):
            eng.use_scorer(ScorerFactory(1.0, 1.0))
            eng.invalidate()
            eng.load_and_round(counts, 2500, 1250, 'constants', want_logfac=eager)
            eng.rounds(2500, 1250, 'constants', 1)
            splits = eng.candidates().copy()
            lmm, total = eng.segment_lmm()
            lf = eng.segment_scores(scores=False, logfac=True)[3]
            results.append((splits, lmm.copy(), total, lf.copy()))
        a, b = results
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and np.array_equal(a[3], b[3])
        from scipy.special import gammaln
        want = np.concatenate([[0.0], np.cumsum(gammaln(counts + 1))])
        assert np.array_equal(a[3], want[a[0]])
        assert a[2] == want[-1]


def test_reducer_accepts_candidates_without_the_ends():
    """the reference's SlidingWindowReducer takes any ascending candidate list and adds 0 and len(counts) to the result
    (sliding_window_reducer.py:22); the fused device loop needs both ends, so such lists take the object route"""
    counts = synth.dnase_like(5000, 3, hotspot_share=0.5)
    cands = np.arange(5001)[7:-9:3]
    factory = ScorerFactory(1.0, 1.0)
    base = ReducerCombiner(NotConstantReducer(), SquareSplitter(factory))
    swr = SlidingWindowReducer(SlidingWindow(100, 50), base)
    got = swr.reduce_candidate_list(counts, cands)
    want = po.sliding_window_round(counts, cands, po.Tables(1, 1.0), 100, 50, 'constants')
    assert np.array_equal(got, want) and got[0] == 0 and got[-1] == 5000


def test_mutated_arrays_are_not_served_from_the_identity_cache():
    """the engine caches the loaded contig / candidates by array identity; an array changed in place between calls (the
    reference recomputes from the array every time) is detected by the sampled fingerprint and uploaded again"""
    counts = synth.dnase_like(20000, 8, hotspot_share=0.5)
    factory = ScorerFactory(1.0, 1.0)
    sq = SquareSplitter(factory)
    cands = np.arange(0, 20001, 40)
    a = sq.split(counts, cands)
    counts[:] = synth.dnase_like(20000, 9, hotspot_share=0.5)          # same object, new content
    b = sq.split(counts, cands)
    o = po.square_split(counts, cands, po.Tables(1, 1.0))
    assert b[0] == o[0] and np.array_equal(b[1], o[1])
    assert a[0] != b[0]
    cands[1:-1] += 1                                                   # same candidate object, new content
    c = sq.split(counts, cands)
    o = po.square_split(counts, cands, po.Tables(1, 1.0))
    assert c[0] == o[0] and np.array_equal(c[1], o[1])


def test_exact_dp_refuses_deep_coverage_clearly():
    """the whole-contig DP indexes the lgamma table with the contig's total count in 32 bits: refused with a clear message
    before any table is built (the default rounds pipeline handles such contigs, test_deep_coverage_total_beyond_int32)"""
    from pasio_b200 import _native
    counts = np.full(3000000, 1000, dtype=np.int64)                    # total 3e9 > 2^31
    with pytest.raises(_native.PasioDeviceError, match='2\\^31'):
        SquareSplitter(ScorerFactory(1.0, 1.0)).split(counts, np.array([0, 1000000, 3000000]))
