"""Test-side debug helper (imports the oracle): step the rounds of a deep-coverage contig one by one against the C oracle."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import _native                                  # noqa: E402
from pasio_b200.log_marginal_likelyhood import ScorerFactory    # noqa: E402
from oracle import c_oracle                                     # noqa: E402

n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
lam = int(sys.argv[2]) if len(sys.argv) > 2 else 110
rs = np.random.RandomState(6)
counts = np.repeat(rs.poisson(lam, n_runs), 50).astype(np.int64)
print('n', len(counts), 'total', counts.sum(), flush=True)
eng = _native.engine()
eng.use_scorer(ScorerFactory(1.0, 1.0))
eng.load(counts)
print('info', eng.info())
Cg = np.concatenate([[0], np.cumsum(counts)])
pos = np.sort(rs.randint(0, len(counts) + 1, 100000)).astype(np.int64)
m = len(pos)
out = np.empty(m, dtype=np.int64)
import ctypes
eng._check(eng.lib.pasio_cumsum_at(eng.ctx, pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), m,
                                   out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
bad = np.flatnonzero(out != Cg[pos])
print('cumsum mismatches', len(bad), (pos[bad[:5]], out[bad[:5]], Cg[pos[bad[:5]]]) if len(bad) else '')
o = c_oracle.FlatOracle(counts, 1.0, 1.0)
cands = np.arange(o.n + 1, dtype=np.int64)
eng.set_candidates(None)
for r in range(8):
    want, _ = o.round(cands, 500, 250, 'constants')
    try:
        n_in, n_out, cells = eng.round(500, 250, 'constants')
    except Exception as e:
        print('round', r + 1, 'FAILED', e)
        break
    got = eng.candidates()
    ok = np.array_equal(got, want)
    print('round', r + 1, n_in, n_out, 'oracle', len(want), 'equal', ok, 'sorted', bool(np.all(np.diff(got) > 0)),
          'range', got.min(), got.max(), flush=True)
    if not ok:
        d = np.setxor1d(got, want)
        print('  xor', len(d), d[:10])
        break
    if len(want) == len(cands):
        break
    cands = want
