"""Generate tests/golden/*.npz from the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):
    python oracle/make_golden.py [--full]
It imports pasio v1.1.3 from /root/reference/src (with oracle/future_shim on the path,
because the `future` package is not installed) and stores inputs + outputs of the
reference's own public API as small fixtures.  Nothing here is used at run time on
the GPU box; the fixtures are.

The reference's tests hold no LogML golden vectors, so these fixtures --
outputs of the reference itself run here -- are what pins oracle/ and the CUDA path.
`tables_sha1` records the host libm/numpy/scipy table bits the fixtures were made
with: on a host whose tables differ (another SIMD dispatch of np.log), bit-level
comparisons against the fixtures are not meaningful and the tests say so.
"""
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
sys.path.insert(0, os.path.join(HERE, 'future_shim'))

import pasio  # noqa: E402  (the reference)
from pasio.splitters import (SquareSplitter, SlidingWindowReducer, RoundReducer,  # noqa: E402
                             NotZeroReducer, NotConstantReducer, ReducerCombiner, NopSplitter,
                             configure_splitter)
from pasio.dto.sliding_window import SlidingWindow  # noqa: E402
from pasio.log_marginal_likelyhood import ScorerFactory  # noqa: E402
from pasio.segmentation import segments_with_scores  # noqa: E402
from pasio.process_bedgraph import split_bedgraph_stream  # noqa: E402

from pasio_b200 import synth  # noqa: E402

assert pasio.__file__.startswith('/root/reference/'), pasio.__file__
OUT = os.path.join(ROOT, 'tests', 'golden')


def tables_sha1():
    import scipy.special
    k = np.arange(1 << 20)
    h = hashlib.sha1()
    h.update(np.log(k + 1.0).tobytes())
    h.update(scipy.special.gammaln(k + 0).tobytes())
    return h.hexdigest()


def exact_case(name, counts, cands, alpha, beta, store, gen=None):
    score, splits = SquareSplitter(ScorerFactory(alpha, beta)).split(counts, cands)
    if gen is None:
        store[name + '.counts'] = counts
    else:   # big inputs are regenerated from their seed: eval(gen, vars(synth))
        store[name + '.gen'] = np.array(gen)
        store[name + '.counts_sha1'] = np.array(hashlib.sha1(counts.tobytes()).hexdigest())
    store[name + '.cands'] = cands
    store[name + '.ab'] = np.array([alpha, beta], dtype=np.float64)
    store[name + '.score'] = np.float64(score)
    store[name + '.splits'] = np.asarray(splits, dtype=np.int64)
    print(name, 'N=%d' % len(cands), 'score=%r' % float(score), 'splits=%d' % len(splits))


def gen_exact(full):
    store = {}
    c = synth.piecewise_poisson(2000, 11)
    allp = np.arange(len(c) + 1)
    exact_case('pp2000_a1b1', c, allp, 1.0, 1.0, store)
    exact_case('pp2000_a3b5', c, allp, 3.0, 5.0, store)
    exact_case('pp2000_a2.5b3', c, allp, 2.5, 3.0, store)
    exact_case('pp2000_a1b0.5', c, allp, 1.0, 0.5, store)
    c = synth.dnase_like(3000, 7, hotspot_share=0.5)
    exact_case('sparse3000_a1b1', c, np.arange(len(c) + 1), 1.0, 1.0, store)
    exact_case('sparse3000_a0.5b1', c, np.arange(len(c) + 1), 0.5, 1.0, store)
    # tests/test_bench_pasio.py:39-48: 100 000 nt, 1001 candidates, sum of counts > 2**20
    c = synth.two_level_poisson(50000, seed=2)
    cands = np.hstack([np.arange(0, len(c), 100), 100000])
    exact_case('bench1001_a1b1', c, cands, 1.0, 1.0, store)
    # spans and counts both past 2**20: the reference takes its unbound (direct ufunc) path
    c = synth.piecewise_poisson(3000000, 5)
    cands = synth.random_candidates(len(c), 1502, 5)
    exact_case('wide1502_a1b1', c, cands, 1.0, 1.0, store, gen='piecewise_poisson(3000000, 5)')
    exact_case('wide1502_a2.5b3', c, cands, 2.5, 3.0, store, gen='piecewise_poisson(3000000, 5)')
    # edge cases: single nt, two nt, all zeros, constant
    exact_case('one_nt', np.array([4]), np.array([0, 1]), 1.0, 1.0, store)
    exact_case('two_nt', np.array([0, 7]), np.array([0, 1, 2]), 1.0, 1.0, store)
    exact_case('zeros500', np.zeros(500, dtype=int), np.arange(501), 1.0, 1.0, store)
    exact_case('const500', np.full(500, 3, dtype=int), np.arange(501), 1.0, 1.0, store)
    store['tables_sha1'] = np.array(tables_sha1())
    np.savez_compressed(os.path.join(OUT, 'exact.npz'), **store)
    if full:
        # BASELINE config 1 in full: n=100 000, all positions (about a minute of CPU)
        store = {}
        c = synth.piecewise_poisson(100000, 0)
        exact_case('config1', c, np.arange(len(c) + 1), 1.0, 1.0, store, gen='piecewise_poisson(100000, 0)')
        del store['config1.cands']   # all positions
        store['tables_sha1'] = np.array(tables_sha1())
        np.savez_compressed(os.path.join(OUT, 'config1.npz'), **store)


def gen_config3():
    """BASELINE config 3 in full through the UNMODIFIED reference: N = 200 000 candidates over 2 Mb
    (SURVEY 8d); several minutes of one core.  Stores the reference's score and splits; the candidates and
    counts are regenerated from their seeds (pasio_b200.synth)."""
    import time
    store = {}
    c = synth.piecewise_poisson(2000000, 1)
    cands = synth.random_candidates(len(c), 200000, 1)
    t0 = time.time()
    score, splits = SquareSplitter(ScorerFactory(1.0, 1.0)).split(c, cands)
    store['config3.gen'] = np.array('piecewise_poisson(2000000, 1)')
    store['config3.cands_gen'] = np.array('random_candidates(2000000, 200000, 1)')
    store['config3.counts_sha1'] = np.array(hashlib.sha1(c.tobytes()).hexdigest())
    store['config3.cands_sha1'] = np.array(hashlib.sha1(cands.tobytes()).hexdigest())
    store['config3.ab'] = np.array([1.0, 1.0])
    store['config3.score'] = np.float64(score)
    store['config3.splits'] = np.asarray(splits, dtype=np.int64)
    store['config3.reference_seconds'] = np.float64(time.time() - t0)
    store['tables_sha1'] = np.array(tables_sha1())
    np.savez_compressed(os.path.join(OUT, 'config3.npz'), **store)
    print('config3 N=%d score=%r splits=%d (%.0f s of the reference)' % (len(cands), float(score), len(splits), time.time() - t0))


def gen_reducers():
    store = {}
    rs = np.random.RandomState(21)
    for k in range(6):
        n = int(rs.randint(5, 400))
        c = (rs.poisson(0.7, n) * (rs.random_sample(n) < 0.5)).astype(int)
        if k == 0:
            c[:] = 0
        inner = np.sort(rs.choice(np.arange(1, n), size=min(n - 1, int(rs.randint(1, n))), replace=False))
        cands = np.hstack([0, inner, n])
        store['r%d.counts' % k] = c
        store['r%d.cands' % k] = cands
        store['r%d.notzero' % k] = np.asarray(NotZeroReducer().reduce_candidate_list(c, cands))
        store['r%d.notconstant' % k] = np.asarray(NotConstantReducer().reduce_candidate_list(c, cands))
    np.savez_compressed(os.path.join(OUT, 'reducers.npz'), **store)


def gen_rounds():
    store = {}
    cases = [
        ('dn60k_c', synth.dnase_like(60000, 13, hotspot_share=0.3), 400, 200, 'constants', 1.0, 1.0),
        ('dn60k_z', synth.dnase_like(60000, 13, hotspot_share=0.3), 400, 200, 'zeros', 1.0, 1.0),
        ('dn20k_n', synth.dnase_like(20000, 14, hotspot_share=0.3), 300, 150, 'none', 1.0, 1.0),
        ('dn60k_real', synth.dnase_like(60000, 15, hotspot_share=0.3), 400, 200, 'constants', 2.5, 3.0),
        ('pp30k_c', synth.piecewise_poisson(30000, 16), 500, 250, 'constants', 1.0, 1.0),
        ('tail_c', synth.piecewise_poisson(1037, 17), 100, 50, 'constants', 1.0, 1.0),
        ('odd_shift', synth.piecewise_poisson(5000, 18), 333, 100, 'constants', 1.0, 1.0),
    ]
    for name, c, wsize, wshift, constraint, alpha, beta in cases:
        factory = ScorerFactory(alpha, beta)
        sq = SquareSplitter(factory)
        base = {'constants': ReducerCombiner(NotConstantReducer(), sq),
                'zeros': ReducerCombiner(NotZeroReducer(), sq), 'none': sq}[constraint]
        swr = SlidingWindowReducer(SlidingWindow(wsize, wshift), base)
        cands = np.arange(len(c) + 1)
        per_round = []
        while True:
            new = swr.reduce_candidate_list(c, cands)
            per_round.append(np.asarray(new, dtype=np.int64))
            if np.array_equal(new, cands):
                break
            cands = new
        final = RoundReducer(swr).reduce_candidate_list(c, np.arange(len(c) + 1))
        assert np.array_equal(final, per_round[-1])
        score, splits = ReducerCombiner(RoundReducer(swr), NopSplitter(factory)).split(c, np.arange(len(c) + 1))
        one_round = RoundReducer(swr, num_rounds=1).reduce_candidate_list(c, np.arange(len(c) + 1))
        assert np.array_equal(one_round, per_round[0])
        store[name + '.counts'] = c
        store[name + '.params'] = np.array([wsize, wshift, alpha, beta], dtype=np.float64)
        store[name + '.constraint'] = np.array(constraint)
        store[name + '.nrounds'] = np.int64(len(per_round))
        for r, arr in enumerate(per_round):
            store[name + '.round%d' % r] = arr
        store[name + '.score'] = np.float64(score)
        store[name + '.splits'] = np.asarray(splits, dtype=np.int64)
        print(name, [len(a) for a in per_round], float(score))
    store['tables_sha1'] = np.array(tables_sha1())
    np.savez_compressed(os.path.join(OUT, 'rounds.npz'), **store)


def gen_pipeline():
    """configure_splitter + segments_with_scores + split_bedgraph text (cli.py:73-84)."""
    store = {}
    cases = [
        ('default300k', synth.dnase_like(300000, 31, hotspot_share=0.3), dict()),
        ('exact4k', synth.piecewise_poisson(4000, 32), dict(algorithm='exact')),
        ('zeros100k', synth.dnase_like(100000, 33, hotspot_share=0.2),
         dict(split_constraints='zeros', window_size=600, window_shift=300)),
        ('real100k', synth.dnase_like(100000, 34, hotspot_share=0.3),
         dict(alpha=0.7, beta=2.0, window_size=1000, window_shift=500)),
        ('rounds2', synth.dnase_like(100000, 35, hotspot_share=0.3),
         dict(num_rounds=2, window_size=1000, window_shift=500)),
    ]
    for name, c, kwargs in cases:
        splitter = configure_splitter(**kwargs)
        segs = list(segments_with_scores(c, splitter))
        store[name + '.counts'] = c
        store[name + '.kwargs'] = np.array(repr(kwargs))
        store[name + '.starts'] = np.array([s.start for s in segs], dtype=np.int64)
        store[name + '.stops'] = np.array([s.stop for s in segs], dtype=np.int64)
        store[name + '.mean'] = np.array([s.mean_count for s in segs], dtype=np.float64)
        store[name + '.lmm'] = np.array([s.log_marginal_likelyhood for s in segs], dtype=np.float64)
        score, splits = splitter.split(c, np.arange(len(c) + 1))
        store[name + '.score'] = np.float64(score)
        print(name, len(segs), float(score))
    # bedgraph text in/out, three output modes, gaps filled and split
    lines = []
    lines += synth.to_bedgraph_lines('chrA', synth.dnase_like(30000, 41, hotspot_share=0.4), chrom_start=100)
    gap = synth.to_bedgraph_lines('chrB', synth.piecewise_poisson(6000, 42))
    lines += [ln for k, ln in enumerate(gap) if k % 7 != 3]          # drop lines -> gaps
    lines += synth.to_bedgraph_lines('chrC', np.array([5], dtype=np.int64), chrom_start=9)
    text = ''.join(lines)
    store['bg.input'] = np.array(text)
    for mode in ['bedgraph', 'bed', 'bedgraph+length+LMM']:
        for gaps in [False, True]:
            out = io.StringIO()
            split_bedgraph_stream(io.StringIO(text), out, configure_splitter(window_size=500, window_shift=250),
                                  split_at_gaps=gaps, output_mode=mode)
            store['bg.%s.%d' % (mode, int(gaps))] = np.array(out.getvalue())
    store['tables_sha1'] = np.array(tables_sha1())
    np.savez_compressed(os.path.join(OUT, 'pipeline.npz'), **store)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    np.seterr(all='ignore')
    if "--config3" in sys.argv:          # only this one (minutes of CPU)
        gen_config3()
        sys.exit(0)
    if "--only-fast" not in sys.argv:
        gen_exact("--full" in sys.argv)
    gen_reducers()
    gen_rounds()
    gen_pipeline()
    print('tables sha1', tables_sha1())
