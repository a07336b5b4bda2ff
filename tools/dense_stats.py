"""Per-round window-DP timing on DENSE data (every position a change point): python tools/dense_stats.py [nt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth, _native
from pasio_b200.log_marginal_likelyhood import ScorerFactory
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 20000000
eng = _native.engine(); eng.use_scorer(ScorerFactory(1.0, 1.0))
eng.load(synth.piecewise_poisson(nt, 7)); eng.set_candidates(None)
for r in range(4):
    eng.timing_reset(True)
    n_in, n_out, cells = eng.round(2500, 1250, 'constants')
    ms = eng.timing()['window_dp'][0]
    c, sk = eng.round_stats()
    print('dense round %d: %d -> %d cands, cells %.4g, skipped %.1f%%, window_dp %.2f ms -> %.3g cells/s' % (r + 1, n_in, n_out, c, 100.0 * sk / max(c, 1), ms, c / ms * 1e3), flush=True)
    if n_in == n_out: break
