"""The oracle (oracle/pasio_oracle.py numpy restatement, oracle/dp_oracle.c) pinned against
fixtures generated from the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import pasio_oracle as po
from oracle import c_oracle
from pasio_b200 import synth

EXACT_CASES = ['pp2000_a1b1', 'pp2000_a3b5', 'pp2000_a2.5b3', 'pp2000_a1b0.5', 'sparse3000_a1b1',
               'sparse3000_a0.5b1', 'bench1001_a1b1', 'wide1502_a1b1', 'wide1502_a2.5b3',
               'one_nt', 'two_nt', 'zeros500', 'const500']


def _tables(ab):
    alpha, beta = float(ab[0]), float(ab[1])
    return po.Tables(po.normalise_alpha(alpha), beta), alpha, beta


@pytest.mark.parametrize('name', EXACT_CASES)
def test_exact_dp_numpy_oracle(golden, name):
    g = golden('exact.npz')
    counts, cands = g.counts(name), g[name + '.cands']
    t, _, _ = _tables(g[name + '.ab'])
    score, splits, _, _ = po.square_split(counts, cands, t)
    g.check_splits(splits, g[name + '.splits'], score, g[name + '.score'], name)


@pytest.mark.parametrize('name', EXACT_CASES)
def test_exact_dp_c_oracle(golden, name):
    g = golden('exact.npz')
    counts, cands = g.counts(name), g[name + '.cands']
    _, alpha, beta = _tables(g[name + '.ab'])
    score, splits, _, _ = c_oracle.FlatOracle(counts, alpha, beta).square_split(cands)
    g.check_splits(splits, g[name + '.splits'], score, g[name + '.score'], name)


def test_reducers(golden):
    g = golden('reducers.npz')
    for k in range(6):
        c, cands = g['r%d.counts' % k], g['r%d.cands' % k]
        assert np.array_equal(po.not_zero(c, cands), g['r%d.notzero' % k])
        assert np.array_equal(po.not_constant(c, cands), g['r%d.notconstant' % k])


ROUND_CASES = ['dn60k_c', 'dn60k_z', 'dn20k_n', 'dn60k_real', 'pp30k_c', 'tail_c', 'odd_shift']


@pytest.mark.parametrize('name', ROUND_CASES)
def test_rounds_numpy_and_c_oracle(golden, name):
    g = golden('rounds.npz')
    counts = g[name + '.counts'].astype(np.int64)
    wsize, wshift, alpha, beta = g[name + '.params']
    wsize, wshift = int(wsize), int(wshift)
    constraint = str(g[name + '.constraint'])
    t = po.Tables(po.normalise_alpha(float(alpha)), float(beta))
    fo = c_oracle.FlatOracle(counts, float(alpha), float(beta))
    cands = np.arange(len(counts) + 1)
    for r in range(int(g[name + '.nrounds'])):
        want = g[name + '.round%d' % r]
        flat, _ = fo.round(cands, wsize, wshift, constraint)
        if r < 2:       # the object-faithful numpy loop is slow; two rounds pin it
            obj = po.sliding_window_round(counts, cands, t, wsize, wshift, constraint)
            assert np.array_equal(obj, flat)
        if g.bit_exact(name):
            assert np.array_equal(flat, want), (name, r)
        cands = flat
    score, splits = po.nop_split(counts, cands, t)
    g.check_splits(splits, g[name + '.splits'], score, g[name + '.score'], name)


def test_default_pipeline_oracle(golden):
    g = golden('pipeline.npz')
    name = 'rounds2'
    counts = g[name + '.counts'].astype(np.int64)
    t = po.Tables(1, 1.0)
    out = po.default_pipeline(counts, t, window_size=1000, window_shift=500, num_rounds=2)
    if g.bit_exact(name):
        assert np.array_equal(out['splits'][:-1], g[name + '.starts'])
        assert np.array_equal(out['splits'][1:], g[name + '.stops'])
        assert np.array_equal(out['mean_counts'], g[name + '.mean'])
        assert np.array_equal(out['lmm'], g[name + '.lmm'])
        assert out['score'] == g[name + '.score']


def test_flat_rounds_match_pipeline_fixture(golden):
    g = golden('pipeline.npz')
    for name, kw in [('default300k', dict(w=2500, s=1250, c='constants', a=1.0, b=1.0)),
                     ('zeros100k', dict(w=600, s=300, c='zeros', a=1.0, b=1.0)),
                     ('real100k', dict(w=1000, s=500, c='constants', a=0.7, b=2.0))]:
        counts = g[name + '.counts'].astype(np.int64)
        fo = c_oracle.FlatOracle(counts, kw['a'], kw['b'])
        cands, sizes, cells = fo.rounds(kw['w'], kw['s'], kw['c'])
        if g.bit_exact(name):
            assert np.array_equal(cands[:-1], g[name + '.starts']), name
            assert np.array_equal(cands[1:], g[name + '.stops']), name


def test_known_answer_log_marginal_likelyhood():
    """reference tests/test_pasio.py:45-65 restated (np.math is gone in numpy 2)."""
    import math

    def exact(counts, alpha, beta):
        fac = 1
        for c in counts:
            fac *= math.factorial(int(c))
        cs, ns = int(sum(counts)), len(counts)
        return np.log((beta ** alpha) * math.gamma(cs + alpha) / (math.gamma(alpha) * fac * ((ns + beta) ** (cs + alpha))))

    for counts, alpha, beta in [([0], 3, 5), ([0, 1], 3, 5), ([4, 0, 1, 3], 5, 2), ([4, 0, 1, 3], 1, 1)]:
        counts = np.array(counts)
        sc = po.Scorer(counts, np.array([0, len(counts)]), po.Tables(alpha, beta))
        assert np.allclose(sc.log_marginal_likelyhoods(), exact(counts, alpha, beta))


def test_config1_fixture_is_consistent(golden):
    """BASELINE config 1 (N=100 001): the fixture's optimal score equals the sum of its segment scores."""
    g = golden('config1.npz')
    counts = g.counts('config1')
    splits = g['config1.splits']
    assert len(splits) == 5123 and counts.sum() == 954615
    sc = po.Scorer(counts, splits, po.Tables(1, 1.0))
    assert abs(np.sum(sc.scores()) - float(g['config1.score'])) <= 1e-9 * abs(float(g['config1.score']))


def test_mt_oracle_equals_single_thread():
    """dp_oracle_mt / round_oracle_mt (OpenMP, used for the full-size comparisons) == the single-threaded loops"""
    from pasio_b200 import synth
    for counts, cands, ab in [(synth.piecewise_poisson(6000, 3), np.arange(6001), (1.0, 1.0)),
                              (synth.dnase_like(8000, 5, hotspot_share=0.4), np.arange(8001), (2.5, 3.0)),
                              (np.zeros(3000, dtype=np.int64), np.arange(3001), (1.0, 1.0))]:
        a = c_oracle.FlatOracle(counts, *ab).square_split(cands)
        b = c_oracle.FlatOracle(counts, *ab, threads=5).square_split(cands)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    counts = synth.dnase_like(400000, 6)
    a = c_oracle.FlatOracle(counts, 1.0, 1.0).rounds(2500, 1250, 'constants')
    b = c_oracle.FlatOracle(counts, 1.0, 1.0, threads=5).rounds(2500, 1250, 'constants')
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2]


def test_config3_oracle_vs_reference_fixture(golden):
    """BASELINE config 3 (N = 200 000 candidates, 2.0e10 cells): the C oracle reproduces the score and the
    6 923 splits of the unmodified reference (tests/golden/config3.npz)."""
    import os
    from pasio_b200 import synth
    g = golden('config3.npz')
    counts = g.counts('config3')
    cands = synth.random_candidates(len(counts), 200000, 1)
    fo = c_oracle.FlatOracle(counts, 1.0, 1.0, threads=max(1, min(32, os.cpu_count() or 1)))
    score, splits, _, _ = fo.square_split(cands)
    g.check_splits(splits, g['config3.splits'], score, g['config3.score'], 'config3')


def test_oracle_text_pipeline_vs_reference_fixture(golden):
    """po.split_bedgraph_text (parser + flat C rounds + reference scoring + '%' formatting) reproduces the reference's
    split_bedgraph_stream output byte for byte: three output modes, gaps filled and split"""
    g = golden('pipeline.npz')
    text = str(g['bg.input'])
    t = po.Tables(1, 1.0)
    for mode in ['bedgraph', 'bed', 'bedgraph+length+LMM']:
        for gaps in [False, True]:
            got = po.split_bedgraph_text(text, t, 500, 250, split_at_gaps=gaps, output_mode=mode)
            if g.bit_exact('oracle text %s %d' % (mode, gaps)):
                assert got == str(g['bg.%s.%d' % (mode, int(gaps))]), (mode, gaps)


def test_chunked_bedgraph_reader_equals_whole_input():
    """the streaming reader (pieces cut at line starts, the contig in progress carried over) yields the same contigs as
    one pass over the whole text -- any piece size, gaps filled or split, blank lines, a contig spanning many pieces"""
    import io
    from pasio_b200 import process_bedgraph as pb, synth
    rs = np.random.RandomState(3)
    lines = []
    for c in range(30):
        n = int(rs.randint(50, 4000)) if c != 7 else 60000
        lines.extend(synth.to_bedgraph_lines('c%d' % (c % 11), synth.dnase_like(n, 100 + c, hotspot_share=0.5),
                                             chrom_start=int(rs.randint(0, 9))))
        if c % 5 == 0:
            lines.append('\n')
    lines = [ln for k, ln in enumerate(lines) if k % 13 != 5]         # gaps
    text = ''.join(lines)
    for gaps in (False, True):
        want = [(c, a.tolist(), s) for c, a, s in po.parse_bedgraph_text(text, gaps)]
        for chunk in (64, 1000, 7777, 1 << 20):
            old = pb.CHUNK_BYTES
            pb.CHUNK_BYTES = chunk
            try:
                got = [(c, a.tolist(), s) for c, a, s in pb.parse_bedgraph_stream(io.StringIO(text), gaps)]
                got_b = [(c, a.tolist(), s) for c, a, s in pb.parse_bedgraph_stream(io.TextIOWrapper(io.BytesIO(text.encode())), gaps)]
            finally:
                pb.CHUNK_BYTES = old
            assert got == want, (gaps, chunk)
            assert got_b == want, (gaps, chunk)
    # a text stream the caller already read a line from: the binary layer is repositioned, nothing is lost
    stream = io.TextIOWrapper(io.BytesIO(text.encode()))
    first = stream.readline()
    rest = [(c, a.tolist(), s) for c, a, s in pb.parse_bedgraph_stream(stream)]
    assert rest == [(c, a.tolist(), s) for c, a, s in po.parse_bedgraph_text(text[len(first):])]
