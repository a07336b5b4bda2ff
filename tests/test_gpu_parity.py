"""GPU parity tests: the CUDA path (through the C ABI, via pasio_b200._native.Engine) against the
oracle and against fixtures generated from the unmodified reference.  Bit-exact for splits,
prefix scores and arg-max indices; 1e-9 relative for summed scores."""
import ctypes
import io

import numpy as np
import pytest

from oracle import pasio_oracle as po
from oracle import c_oracle
from pasio_b200 import synth, _native
from pasio_b200.log_marginal_likelyhood import ScorerFactory

pytestmark = pytest.mark.gpu

_factories = {}


def factory(alpha, beta):
    key = (alpha, beta)
    if key not in _factories:
        _factories[key] = ScorerFactory(alpha, beta)
    return _factories[key]


@pytest.fixture(scope='module')
def eng():
    return _native.engine()


def gpu_exact(eng, counts, cands, alpha, beta, arrays=True):
    eng.use_scorer(factory(alpha, beta))
    eng.load(counts)
    all_pos = len(cands) == len(counts) + 1
    eng.set_candidates(None if all_pos else cands)
    return eng.square_split(want_arrays=arrays)


EXACT_CASES = ['pp2000_a1b1', 'pp2000_a3b5', 'pp2000_a2.5b3', 'pp2000_a1b0.5', 'sparse3000_a1b1',
               'sparse3000_a0.5b1', 'bench1001_a1b1', 'wide1502_a1b1', 'wide1502_a2.5b3',
               'one_nt', 'two_nt', 'zeros500', 'const500']


@pytest.mark.parametrize('name', EXACT_CASES)
def test_exact_dp_vs_reference_fixture(eng, golden, name):
    g = golden('exact.npz')
    counts, cands = g.counts(name), g[name + '.cands'].astype(np.int64)
    alpha, beta = [float(x) for x in g[name + '.ab']]
    score, splits, P, prev = gpu_exact(eng, counts, cands, alpha, beta)
    g.check_splits(splits, g[name + '.splits'], score, g[name + '.score'], name)
    # and bit for bit against the oracle's DP arrays on this host
    o_score, o_splits, o_P, o_prev = c_oracle.FlatOracle(counts, alpha, beta).square_split(cands)
    assert np.array_equal(P, o_P) and np.array_equal(prev, o_prev)
    assert score == o_score and np.array_equal(splits, o_splits)


@pytest.mark.parametrize('n', [1, 2, 31, 32, 33, 34, 63, 64, 65, 127, 128, 129, 130, 255, 256, 257, 258, 300, 1000, 4100])
@pytest.mark.parametrize('ab', [(1.0, 1.0), (2.5, 3.0)])
def test_exact_dp_block_edges(eng, n, ab):
    """candidate counts around the 32-row / 128-row / 256-column tile edges"""
    rs = np.random.RandomState(n)
    counts = (rs.poisson(3, n) * (rs.random_sample(n) < 0.6)).astype(np.int64)
    cands = np.arange(n + 1, dtype=np.int64)
    score, splits, P, prev = gpu_exact(eng, counts, cands, *ab)
    o_score, o_splits, o_P, o_prev = c_oracle.FlatOracle(counts, *ab).square_split(cands)
    assert np.array_equal(P, o_P)
    assert np.array_equal(prev, o_prev)
    assert np.array_equal(splits, o_splits) and score == o_score


def test_exact_dp_sparse_candidates_and_ties(eng):
    counts = synth.dnase_like(200000, 77, hotspot_share=0.3)
    cands = synth.random_candidates(len(counts), 3000, 78)
    for ab in [(1.0, 1.0), (0.5, 2.0), (4.0, 0.25)]:
        score, splits, P, prev = gpu_exact(eng, counts, cands, *ab)
        o_score, o_splits, o_P, o_prev = c_oracle.FlatOracle(counts, *ab).square_split(cands)
        assert np.array_equal(P, o_P) and np.array_equal(prev, o_prev)
        assert np.array_equal(splits, o_splits)


def test_suffix_rows_vs_numpy_oracle(eng):
    counts = synth.piecewise_poisson(5000, 3)
    for alpha, beta, cands in [(1.0, 1.0, np.arange(5001)), (2.5, 3.0, synth.random_candidates(5000, 400, 1))]:
        f = factory(alpha, beta)
        eng.use_scorer(f)
        eng.load(counts)
        eng.set_candidates(None if len(cands) == 5001 else cands)
        sc = po.Scorer(counts, cands, po.Tables(po.normalise_alpha(alpha), beta))
        for stop in [1, 2, 33, len(cands) // 2, len(cands) - 1]:
            assert np.array_equal(eng.suffix_scores(stop), sc.row(stop)), (alpha, stop)
        assert np.array_equal(eng.cumsum_at_candidates(), sc.cumsum)


ROUND_CASES = ['dn60k_c', 'dn60k_z', 'dn20k_n', 'dn60k_real', 'pp30k_c', 'tail_c', 'odd_shift']


@pytest.mark.parametrize('name', ROUND_CASES)
def test_rounds_vs_reference_fixture(eng, golden, name):
    g = golden('rounds.npz')
    counts = g[name + '.counts'].astype(np.int64)
    wsize, wshift, alpha, beta = g[name + '.params']
    wsize, wshift, alpha, beta = int(wsize), int(wshift), float(alpha), float(beta)
    constraint = str(g[name + '.constraint'])
    fo = c_oracle.FlatOracle(counts, alpha, beta)
    eng.use_scorer(factory(alpha, beta))
    eng.load(counts)
    eng.set_candidates(None)
    cands = np.arange(len(counts) + 1, dtype=np.int64)
    nrounds = int(g[name + '.nrounds'])
    for r in range(nrounds):
        n_in, n_out, cells = eng.round(wsize, wshift, constraint)
        got = eng.candidates()
        want, o_cells = fo.round(cands, wsize, wshift, constraint)
        assert np.array_equal(got, want), (name, r)
        assert (n_in, n_out, cells) == (len(cands), len(want), o_cells)
        if g.bit_exact(name):
            assert np.array_equal(got, g[name + '.round%d' % r]), (name, r)
        cands = got
    # the whole loop in one call
    eng.set_candidates(None)
    sizes, final, _ = eng.rounds(wsize, wshift, constraint)
    assert len(sizes) == nrounds and final == len(cands)
    assert np.array_equal(eng.candidates(), cands)
    # explicit starting candidates (round 2 input) give round 2 output
    if nrounds >= 2:
        start = fo.round(np.arange(len(counts) + 1), wsize, wshift, constraint)[0]
        eng.set_candidates(start)
        eng.round(wsize, wshift, constraint)
        assert np.array_equal(eng.candidates(), fo.round(start, wsize, wshift, constraint)[0])
    # final scoring (NopSplitter) against the fixture
    eng.set_candidates(cands)
    scores, segc, means, logfac = eng.segment_scores(scores=True, counts=True, means=True, logfac=True)
    sc = po.Scorer(counts, cands, po.Tables(po.normalise_alpha(alpha), beta))
    assert np.array_equal(scores, sc.scores())
    assert np.array_equal(means, sc.mean_counts())
    assert np.array_equal(segc, np.diff(sc.cumsum))
    assert np.array_equal(logfac, sc.logfac_cumsum)            # sequential sum, like np.cumsum (csrc/logfac_exact.cu)
    g.check_splits(cands, g[name + '.splits'], np.sum(scores), g[name + '.score'], name)


def test_max_rounds_and_resume(eng):
    counts = synth.dnase_like(60000, 13, hotspot_share=0.3)
    fo = c_oracle.FlatOracle(counts, 1.0, 1.0)
    eng.use_scorer(factory(1.0, 1.0))
    eng.load(counts)
    for limit in [1, 2, 3]:
        eng.set_candidates(None)
        sizes, final, _ = eng.rounds(400, 200, 'constants', num_rounds=limit)
        want, o_sizes, _ = fo.rounds(400, 200, 'constants', num_rounds=limit)
        assert np.array_equal(eng.candidates(), want) and len(sizes) == limit


def test_batch_equals_per_contig(eng):
    lens = [1, 2, 700, 5000, 1249, 1251, 2501, 20000, 3]
    parts = [synth.dnase_like(n, 100 + k, hotspot_share=0.5) for k, n in enumerate(lens)]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    counts = np.concatenate(parts)
    eng.use_scorer(factory(1.0, 1.0))
    for constraint in ['constants', 'zeros', 'none']:
        eng.load(counts, offsets=offsets)
        eng.set_candidates(None)
        eng.rounds(300, 150, constraint)
        got = eng.candidates()
        want = []
        for k, part in enumerate(parts):
            c, _, _ = c_oracle.FlatOracle(part, 1.0, 1.0).rounds(300, 150, constraint)
            want.append(c + offsets[k])
        want = np.unique(np.concatenate(want))
        assert np.array_equal(got, want), constraint
    eng.load(counts)        # leave the engine in single-contig mode


def test_rle_load_equals_dense(eng):
    counts = synth.dnase_like(300000, 5, hotspot_share=0.2)
    change = np.flatnonzero(counts[1:] != counts[:-1]) + 1
    starts = np.concatenate([[0], change, [len(counts)]]).astype(np.int64)
    values = counts[starts[:-1]]
    eng.use_scorer(factory(1.0, 1.0))
    eng.load_rle(starts, values)
    assert eng.info()[:2] == (len(counts), int(counts.sum()))
    eng.rounds(500, 250, 'constants')
    a = eng.candidates()
    eng.load(counts)
    eng.set_candidates(None)
    eng.rounds(500, 250, 'constants')
    assert np.array_equal(a, eng.candidates())


def test_error_behaviour(eng):
    eng.use_scorer(factory(1.0, 1.0))
    bad = np.array([1, 2, -1, 4], dtype=np.int64)
    with pytest.raises(AssertionError):
        eng.load(bad)
    good = np.array([1, 2, 0, 4, 4, 4], dtype=np.int64)
    eng.load(good)
    for cands in [[1, 3, 6], [0, 3, 5], [0, 3, 3, 6], [0, 4, 2, 6]]:
        with pytest.raises(AssertionError):
            eng.set_candidates(np.array(cands, dtype=np.int64))
    eng.set_candidates(np.array([0, 3, 6], dtype=np.int64))
    assert eng.candidate_count() == 3
    # a window that cannot fit one CTA's shared memory is refused, not silently mis-computed
    big = synth.piecewise_poisson(40000, 1) + 1
    eng.load(big)
    eng.set_candidates(None)
    with pytest.raises(_native.PasioDeviceError):
        eng.round(30000, 15000, 'none')


def test_config1_full_vs_reference_fixture(eng, golden):
    """BASELINE config 1 in full: n=100 000 nt, all positions candidates (5.0e9 cells)."""
    g = golden('config1.npz')
    counts = g.counts('config1')
    score, splits = gpu_exact(eng, counts, np.arange(len(counts) + 1), 1.0, 1.0, arrays=False)
    g.check_splits(splits, g['config1.splits'], score, g['config1.score'], 'config1')
    # size-independent properties: the optimum equals the sum of its segment scores, and is a fixed point
    eng.set_candidates(splits)
    scores = eng.segment_scores(scores=True)[0]
    assert abs(np.sum(scores) - score) <= 1e-9 * abs(score)
    score2, splits2 = eng.square_split()
    assert np.array_equal(splits2, splits)


def _oracle_threads():
    import os
    return max(1, min(32, os.cpu_count() or 1))


def test_config1_full_dp_arrays_vs_oracle(eng):
    """config 1 at its stated shape: every prefix score and arg-max of the 100 001 rows equals the C oracle
    (the pruned kernel skips > 95 % of the cells: a wrong bound would show here)."""
    counts = synth.piecewise_poisson(100000, 0)
    cands = np.arange(len(counts) + 1, dtype=np.int64)
    score, splits, P, prev = gpu_exact(eng, counts, cands, 1.0, 1.0)
    cells, skipped = eng.round_stats()
    assert cells == len(cands) * (len(cands) - 1) // 2 and 0 < skipped < cells
    o_score, o_splits, o_P, o_prev = c_oracle.FlatOracle(counts, 1.0, 1.0, threads=_oracle_threads()).square_split(cands)
    assert np.array_equal(P, o_P) and np.array_equal(prev, o_prev)
    assert score == o_score and np.array_equal(splits, o_splits)


def test_config3_full_vs_reference_fixture_and_oracle(eng, golden):
    """BASELINE config 3 at its stated shape (N = 200 000 candidates over 2 Mb, 2.0e10 cells): splits and score
    equal the UNMODIFIED reference's (tests/golden/config3.npz, 827 s of the reference), and the complete P / prev
    arrays equal the C oracle's."""
    import hashlib
    g = golden('config3.npz')
    counts = g.counts('config3')
    cands = synth.random_candidates(len(counts), 200000, 1)
    assert hashlib.sha1(cands.tobytes()).hexdigest() == str(g['config3.cands_sha1'])
    score, splits, P, prev = gpu_exact(eng, counts, cands, 1.0, 1.0)
    g.check_splits(splits, g['config3.splits'], score, g['config3.score'], 'config3')
    o_score, o_splits, o_P, o_prev = c_oracle.FlatOracle(counts, 1.0, 1.0, threads=_oracle_threads()).square_split(cands)
    assert np.array_equal(P, o_P) and np.array_equal(prev, o_prev)
    assert score == o_score and np.array_equal(splits, o_splits)
    # the un-pruned kernel (every cell evaluated) gives the same arrays
    eng.set_tuning('exact_prune', 0)
    try:
        score0, splits0, P0, prev0 = gpu_exact(eng, counts, cands, 1.0, 1.0)
    finally:
        eng.set_tuning('exact_prune', 1)
    assert np.array_equal(P0, P) and np.array_equal(prev0, prev)


PRUNED_EXACT_CASES = {
    'zeros': lambda: (np.zeros(20000, dtype=np.int64), None),                      # arg-max is always column 0 (far)
    'const': lambda: (np.full(20000, 3, dtype=np.int64), None),
    'sparse_ties': lambda: (synth.dnase_like(30000, 5, hotspot_share=0.3), None),  # exact ties in the arg-max
    'dense': lambda: (synth.two_level_poisson(12000, seed=3), None),
    'pp_cands': lambda: (synth.piecewise_poisson(300000, 9), synth.random_candidates(300000, 25000, 9)),
    'long_segments': lambda: (np.repeat(np.random.RandomState(4).poisson(20, 12), 2500).astype(np.int64)
                              + np.random.RandomState(5).poisson(3, 30000), None),
    # every candidate is a split point (levels far apart, candidates exactly at the steps): the dependency chains inside a
    # 32-row block are long (the opposite of the other cases, where the arg-max is far behind the block)
    'all_split': lambda: (np.repeat(np.arange(9000) % 2 * 400 + np.arange(9000) % 7 * 60, 5).astype(np.int64),
                          np.arange(0, 45001, 5, dtype=np.int64)),
}


@pytest.mark.parametrize('case', sorted(PRUNED_EXACT_CASES))
@pytest.mark.parametrize('ab', [(1.0, 1.0), (2.5, 3.0)])
def test_exact_pruned_vs_oracle_and_unpruned(eng, case, ab):
    """the pruned whole-contig DP (bounded far columns, both lags) against the oracle and against the kernel
    that evaluates every cell: P and prev bit for bit, on data where the far columns win, tie, or never matter"""
    counts, cands = PRUNED_EXACT_CASES[case]()
    if cands is None:
        cands = np.arange(len(counts) + 1, dtype=np.int64)
    o_score, o_splits, o_P, o_prev = c_oracle.FlatOracle(counts, *ab, threads=_oracle_threads()).square_split(cands)
    try:
        for prune, lag, ring, nblock in [(1, 3, 1, 1), (1, 4, 1, 1), (1, 3, 0, 1), (0, 3, 1, 1), (1, 3, 1, 0), (1, 4, 0, 0), (1, 4, 1, 2), (1, 4, 0, 2), (1, 5, 1, 3), (1, 5, 0, 2)]:
            eng.set_tuning('exact_prune', prune)
            eng.set_tuning('exact_lag', lag)
            eng.set_tuning('exact_ring', ring)       # 1: self scores in a ring of 64 slabs (default), 0: one slab per block
            eng.set_tuning('exact_nblock', nblock)   # 1 / 2: the band's first column block(s) on worker CTAs, 0: swept by the diagonal
            score, splits, P, prev = gpu_exact(eng, counts, cands, *ab)
            assert np.array_equal(P, o_P), (case, prune, lag)
            assert np.array_equal(prev, o_prev), (case, prune, lag)
            assert score == o_score and np.array_equal(splits, o_splits)
    finally:
        eng.set_tuning('exact_prune', 1)
        eng.set_tuning('exact_lag', 5)
        eng.set_tuning('exact_ring', 1)
        eng.set_tuning('exact_nblock', 3)


def test_config3_prefix_property(eng):
    """BASELINE config 3: N=200 000 candidates over 2 Mb.  The first K rows of the DP depend only on
    the first K candidates, so the oracle pins a prefix of the full-size run bit for bit."""
    counts = synth.piecewise_poisson(2000000, 1)
    cands = synth.random_candidates(len(counts), 200000, 1)
    score, splits, P, prev = gpu_exact(eng, counts, cands, 1.0, 1.0)
    K = 20000
    sub_counts = counts[:cands[K - 1]]
    _, _, o_P, o_prev = c_oracle.FlatOracle(sub_counts, 1.0, 1.0).square_split(cands[:K])
    assert np.array_equal(P[:K], o_P) and np.array_equal(prev[:K], o_prev)
    assert splits[0] == 0 and splits[-1] == len(counts) and np.all(np.diff(splits) > 0)
    assert np.all(np.isin(splits, cands))
    eng.set_candidates(splits)
    scores = eng.segment_scores(scores=True)[0]
    assert abs(np.sum(scores) - score) <= 1e-9 * abs(score)


def test_default_pipeline_5mb_vs_oracle(eng):
    """default flags (2500/1250, constants) on a 5 Mb DNase-like contig, every round checked"""
    counts = synth.dnase_like(5000000, 0)
    fo = c_oracle.FlatOracle(counts, 1.0, 1.0)
    eng.use_scorer(factory(1.0, 1.0))
    eng.load(counts)
    eng.set_candidates(None)
    cands = np.arange(len(counts) + 1, dtype=np.int64)
    for r in range(20):
        n_in, n_out, cells = eng.round(2500, 1250, 'constants')
        want, o_cells = fo.round(cands, 2500, 1250, 'constants')
        got = eng.candidates()
        assert np.array_equal(got, want), r
        assert cells == o_cells
        if len(want) == len(cands):
            break
        cands = want
    assert r >= 2


@pytest.mark.parametrize('case', ['dense_none', 'dense_real', 'ties_zero_one', 'big_counts'])
def test_pruned_far_columns_are_exact(eng, case):
    """full-size windows (2500/1250) where the branch-and-bound skips most far columns: every round must
    still equal the oracle, and the kernel must really have skipped cells (so the path under test ran)"""
    rs = np.random.RandomState(5)
    if case == 'dense_none':
        counts, ab, constraint = synth.piecewise_poisson(60000, 21), (1.0, 1.0), 'none'
    elif case == 'dense_real':
        counts, ab, constraint = synth.piecewise_poisson(150000, 22), (0.37, 2.5), 'constants'
    elif case == 'ties_zero_one':
        counts, ab, constraint = (rs.random_sample(400000) < 0.3).astype(np.int64), (1.0, 1.0), 'constants'
    else:
        counts, ab, constraint = (synth.piecewise_poisson(120000, 23) * 500 + rs.poisson(3, 120000)).astype(np.int64), (2.0, 0.5), 'constants'
    fo = c_oracle.FlatOracle(counts, *ab)
    eng.use_scorer(factory(*ab))
    eng.load(counts)
    eng.set_candidates(None)
    cands = np.arange(len(counts) + 1, dtype=np.int64)
    skipped_total = 0
    for r in range(6):
        eng.round(2500, 1250, constraint)
        got = eng.candidates()
        want, o_cells = fo.round(cands, 2500, 1250, constraint)
        cells, skipped = eng.round_stats()
        assert np.array_equal(got, want), (case, r)
        assert cells == o_cells and 0 <= skipped < cells
        skipped_total += skipped
        if len(want) == len(cands):
            break
        cands = want
    assert skipped_total > 0


@pytest.mark.parametrize('wsize', [159, 160, 511, 512, 513])
@pytest.mark.parametrize('explicit', [False, True])
def test_window_size_class_edges(eng, wsize, explicit):
    """windows of exactly 160 / 161 / 512 / 513 / 514 candidates sit on the edges between the warp-per-window
    kernels (<= 160, <= 512 candidates) and the CTA-per-window kernel; all three must give the oracle's rounds"""
    rs = np.random.RandomState(wsize)
    counts = rs.poisson(3.0, 9000).astype(np.int64)
    fo = c_oracle.FlatOracle(counts, 1.0, 1.0)
    eng.use_scorer(factory(1.0, 1.0))
    eng.load(counts)
    if explicit:
        cands = np.concatenate([[0], np.flatnonzero(rs.random_sample(len(counts) - 1) < 0.7) + 1, [len(counts)]]).astype(np.int64)
        eng.set_candidates(cands)
    else:
        cands = np.arange(len(counts) + 1, dtype=np.int64)
        eng.set_candidates(None)
    for r in range(3):
        eng.round(wsize, wsize // 2, 'none')
        got = eng.candidates()
        want, o_cells = fo.round(cands, wsize, wsize // 2, 'none')
        assert np.array_equal(got, want), (wsize, explicit, r)
        assert eng.round_stats()[0] == o_cells
        if len(want) == len(cands):
            break
        cands = want


@pytest.mark.parametrize('wsize,wshift', [(600, 7), (600, 199), (600, 300), (600, 301), (600, 600), (600, 1000), (2500, 100)])
def test_window_geometry_strides(eng, wsize, wshift):
    """window skipping (a phase-2 window all of whose candidates already survived is not computed) for shifts that
    are not half a window: many windows per candidate, no overlap at all, gaps between windows"""
    rs = np.random.RandomState(wsize + wshift)
    counts = synth.dnase_like(24000 if wshift < 100 else 60000, wsize + wshift, hotspot_share=0.4)
    fo = c_oracle.FlatOracle(counts, 1.0, 1.0)
    eng.use_scorer(factory(1.0, 1.0))
    eng.load(counts)
    cands = np.concatenate([[0], np.flatnonzero(rs.random_sample(len(counts) - 1) < 0.5) + 1, [len(counts)]]).astype(np.int64)
    eng.set_candidates(cands)
    for r in range(4):
        eng.round(wsize, wshift, 'none')
        got = eng.candidates()
        want, o_cells = fo.round(cands, wsize, wshift, 'none')
        assert np.array_equal(got, want), (wsize, wshift, r)
        assert eng.round_stats()[0] == o_cells
        if len(want) == len(cands):
            break
        cands = want


@pytest.mark.parametrize('n,constraint', [(70000, 'constants'), (5000000, 'constants'), (40000000, 'constants'),
                                          (300000, 'none'), (300000, 'zeros')])
def test_load_and_round_equals_load_then_round(eng, n, constraint):
    """pasio_contig_load_round (chunked upload overlapped with the scan and the first round; 40 Mb = two chunks)
    gives the state and the candidates of pasio_contig_load + pasio_round, and the rounds after it agree too"""
    counts = synth.dnase_like(n, 77, hotspot_share=0.2)
    eng.use_scorer(factory(1.0, 1.0))
    eng.invalidate()
    eng.load(counts)
    eng.set_candidates(None)
    a_in, a_out, a_cells = eng.round(2500, 1250, constraint)
    want1 = eng.candidates().copy()
    sizes_a, final_a, _ = eng.rounds(2500, 1250, constraint)
    want = eng.candidates().copy()
    total_a = eng.info()[1]
    eng.invalidate()
    first = eng.load_and_round(counts, 2500, 1250, constraint)
    assert first == (a_in, a_out, a_cells)
    assert eng.info() == (n, total_a, 1)
    assert np.array_equal(eng.candidates(), want1)
    sizes_b, final_b, _ = eng.rounds(2500, 1250, constraint, first=first)
    assert sizes_b == [a_in] + sizes_a and final_b == final_a
    assert np.array_equal(eng.candidates(), want)


def test_load_and_round_narrowed_upload(eng):
    """counts cross PCIe as uint8 / uint16 / int32 where they fit (host threads pack them, kernels widen them): same state
    and same first round as with plain copies, an eighth of the bytes on the link; with the limits lowered (test switch:
    uint8 below 2^3, uint16 below 2^6, int32 below 2^9) all four kinds of slices occur in one run"""
    n = 40000003                                     # (a ragged last slice: it goes up as it is)
    counts = synth.dnase_like(n, 78, hotspot_share=0.2)
    counts[counts >= 8] = 7                          # (so that whole slices fit 3 bits: uint8 under the test switch ...)
    counts[12000000:15000000:499] = 40               # (... some need uint16 under the switch ...)
    counts[3000000:9000000:997] = 300                # (... some int32; uint16 by default ...)
    counts[30000001] = 70000                         # (... and one is plain under the switch, int32 by default)
    eng.use_scorer(factory(1.0, 1.0))
    got = {}
    try:
        for mode in (0, 1, 3):
            eng.set_tuning('upload_narrow', mode)
            eng.invalidate()
            first = eng.load_and_round(counts, 2500, 1250, 'constants')
            got[mode] = (first, eng.info(), eng.candidates().copy(), eng.upload_stats())
    finally:
        eng.set_tuning('upload_narrow', 1)
    for mode in (1, 3):
        assert got[mode][0] == got[0][0] and got[mode][1] == got[0][1]
        assert np.array_equal(got[mode][2], got[0][2])
    assert got[0][1][1] == int(counts.sum())
    assert got[0][3] == 8 * n
    assert n <= got[1][3] < n + 24 * (1 << 19)                 # uint8 but for a dozen uint16 slices, one int32, the ragged last one plain
    assert got[1][3] < got[3][3] < 8 * n                       # some slices narrowed, some not
    # the dense profile on the device is what was sent: scores of a few fixed segments need the true prefix sums
    want = np.concatenate([[0], np.cumsum(counts)])[[0, 5, n // 3, n - 7, n]]
    eng.set_candidates(np.array([0, 5, n // 3, n - 7, n], dtype=np.int64))
    assert np.array_equal(eng.cumsum_at_candidates(), want)
    # the plain load (batches of contigs) takes the same route
    try:
        for mode, lo, hi in [(1, n, n + 24 * (1 << 19)), (3, n + 1, 8 * n - 1), (0, 8 * n, 8 * n)]:
            eng.set_tuning('upload_narrow', mode)
            eng.invalidate()
            eng.load(counts)
            assert lo <= eng.upload_stats() <= hi
            assert eng.info() == (n, int(counts.sum()), 1)
            eng.set_candidates(np.array([0, 5, n // 3, n - 7, n], dtype=np.int64))
            assert np.array_equal(eng.cumsum_at_candidates(), want)
    finally:
        eng.set_tuning('upload_narrow', 1)


def test_load_and_round_asserts_and_table_growth(eng):
    """negative counts still raise; a first round that needs longer tables falls back to load + grow + round"""
    eng.use_scorer(ScorerFactory(1.0, 1.0))            # fresh factory: tables at their initial 2^20 entries
    bad = np.ones(5000, dtype=np.int64)
    bad[1234] = -1
    with pytest.raises(AssertionError):
        eng.load_and_round(bad, 2500, 1250, 'constants')
    # a LARGE negative count: prefix sums stop being monotone, so no window may be processed after the scan saw it
    # (table indices would be far out of range); the context must stay usable afterwards
    for n, at in [(300000, 150000), (40000000, 39000000)]:
        big = synth.dnase_like(n, 3, hotspot_share=0.3)
        big[at] = -10 ** 9
        with pytest.raises(AssertionError):
            eng.load_and_round(big, 2500, 1250, 'constants')
        with pytest.raises(AssertionError):
            eng.load(big)
    ok = synth.dnase_like(50000, 4, hotspot_share=0.3)
    first_ok = eng.load_and_round(ok, 2500, 1250, 'constants')
    want_ok, cells_ok = c_oracle.FlatOracle(ok, 1.0, 1.0).round(np.arange(len(ok) + 1, dtype=np.int64), 2500, 1250, 'constants')
    assert np.array_equal(eng.candidates(), want_ok) and first_ok[2] == cells_ok
    deep = np.repeat(np.random.RandomState(3).poisson(900, 4000), 25).astype(np.int64)      # window counts > 2^20
    fo = c_oracle.FlatOracle(deep, 1.0, 1.0)
    first = eng.load_and_round(deep, 2500, 1250, 'constants')
    want, cells = fo.round(np.arange(len(deep) + 1, dtype=np.int64), 2500, 1250, 'constants')
    assert np.array_equal(eng.candidates(), want) and first == (len(deep) + 1, len(want), cells)


@pytest.mark.parametrize('n,wsize', [(3000, 50), (40000, 300), (700000, 2500), (6000000, 2500)])
def test_segment_scores_sum_is_numpy_sum(eng, n, wsize):
    """pasio_segment_scores_sum restates np.sum's pairwise order on the device: same float64 as np.sum(scores)"""
    counts = synth.dnase_like(n, n % 97, hotspot_share=0.3)
    eng.use_scorer(factory(1.0, 1.0))
    eng.load(counts)
    eng.set_candidates(None)
    eng.rounds(wsize, wsize // 2, 'constants', 2)
    scores, _, _, _ = eng.segment_scores(scores=True)
    assert eng.segment_scores_sum() == np.sum(scores)
    eng.set_candidates(np.array([0, n], dtype=np.int64))                 # one segment
    scores, _, _, _ = eng.segment_scores(scores=True)
    assert eng.segment_scores_sum() == np.sum(scores)


def test_timing_hooks(eng):
    counts = synth.dnase_like(100000, 9, hotspot_share=0.3)
    eng.use_scorer(factory(1.0, 1.0))
    eng.timing_reset(True)
    eng.load(counts)
    eng.set_candidates(None)
    eng.rounds(500, 250, 'constants')
    t = eng.timing()
    eng.timing_reset(False)
    assert t['scan'][1] == 1 and t['scan'][0] > 0
    assert t['window_dp'][1] >= 2 and t['window_dp'][0] > 0


def test_randomised_rounds_fuzz():
    """tests/fuzz_parity.py for a few seconds: random data kinds, alpha/beta, window geometry and constraints,
    every round compared with the oracle (the longer runs are recorded in profiles/)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'tests', 'fuzz_parity.py'), '20', '7'],
                         capture_output=True, text=True, cwd=root)
    assert out.returncode == 0 and 'fuzz ok' in out.stdout, out.stdout + out.stderr
