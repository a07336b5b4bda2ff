"""Determinism stress: the same contigs through load + rounds over and over, interleaved, in one process.
Any difference between repetitions (or an error) is reported.  Usage: python tools/stress_rounds.py [iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import _native, synth                           # noqa: E402
from pasio_b200.log_marginal_likelyhood import ScorerFactory    # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rs = np.random.RandomState(6)
cases = {
    'deep20M': (np.repeat(rs.poisson(110, 400000), 50).astype(np.int64), 500, 250),
    'deep2M': (np.repeat(rs.poisson(110, 40000), 50).astype(np.int64), 500, 250),
    'dnase8M': (synth.dnase_like(8000000, 3), 2500, 1250),
    'pw1M': (synth.piecewise_poisson(1000000, 4), 300, 100),
}
eng = _native.engine()
first = {}
t0 = time.time()
for it in range(iters):
    for name, (counts, size, shift) in cases.items():
        if name == 'deep20M' and it % 5:
            continue
        eng.use_scorer(ScorerFactory(1, 1))      # fresh factory: tables re-uploaded at 2^20 and re-grown
        eng.invalidate()
        eng.load(counts)
        eng.set_candidates(None)
        try:
            sizes, final, cells = eng.rounds(size, shift, 'constants')
            got = eng.candidates().copy()
        except Exception as e:           # noqa: BLE001
            print('iter', it, name, 'ERROR', e, flush=True)
            continue
        if name not in first:
            first[name] = (sizes, got)
            print('iter', it, name, 'sizes', sizes, 'final', final, flush=True)
        elif sizes != first[name][0] or not np.array_equal(got, first[name][1]):
            print('iter', it, name, 'MISMATCH sizes', sizes, 'vs', first[name][0], flush=True)
print('done', iters, 'iterations in %.1f s' % (time.time() - t0))
