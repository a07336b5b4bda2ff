"""BASELINE config 2 at its stated shape against the oracle: the chr1-sized bench contig (248 956 422 nt, seed 1000), default
flags, EVERY round of the device pipeline compared with oracle/dp_oracle.c's round (windows spread over the host cores with
OpenMP), then the final splits / scores / means with the numpy restatement of the scorer.  Writes one JSON line.

    python tools/parity_chr1.py [--nt N] > gpurun_out/parity_chr1.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth, _native                            # noqa: E402
from pasio_b200.log_marginal_likelyhood import ScorerFactory     # noqa: E402
from oracle import c_oracle, pasio_oracle as po                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nt', type=int, default=248956422)
args = ap.parse_args()
threads = os.cpu_count() or 1
t0 = time.time()
counts = synth.dnase_like(args.nt, seed=1000)
fo = c_oracle.FlatOracle(counts, 1.0, 1.0, threads=threads)
eng = _native.engine()
eng.use_scorer(ScorerFactory(1.0, 1.0))
eng.load(counts)
eng.set_candidates(None)
cands = np.arange(len(counts) + 1, dtype=np.int64)
rounds, all_equal = [], True
while True:
    n_in, n_out, cells = eng.round(2500, 1250, 'constants')
    got = eng.candidates()
    t1 = time.time()
    want, o_cells = fo.round(cands, 2500, 1250, 'constants')
    equal = bool(np.array_equal(got, want)) and cells == o_cells and n_in == len(cands) and n_out == len(want)
    rounds.append({'round': len(rounds) + 1, 'candidates_in': int(n_in), 'candidates_out': int(n_out), 'cells': int(cells),
                   'cells_skipped_by_bound': int(eng.round_stats()[1]), 'equal_to_oracle': equal, 'oracle_seconds': round(time.time() - t1, 2)})
    print(rounds[-1], file=sys.stderr, flush=True)
    all_equal = all_equal and equal
    if n_in == n_out or not equal:
        break
    cands = np.array(got, dtype=np.int64)
scores, segc, means, _ = eng.segment_scores(scores=True, counts=True, means=True)
total = eng.segment_scores_sum()
lmm, sum_logfac = eng.segment_lmm()
sc = po.Scorer(counts, np.asarray(cands if not all_equal else eng.candidates()), po.Tables(1, 1.0))
final = {'scores_equal': bool(np.array_equal(scores, sc.scores())), 'means_equal': bool(np.array_equal(means, sc.mean_counts())),
         'total_equals_np_sum': bool(total == np.sum(sc.scores())),
         'lmm_equal': bool(np.array_equal(lmm, sc.log_marginal_likelyhoods())), 'sum_logfac_equal': bool(sum_logfac == sc.total_sum_logfac())}
print(json.dumps({'workload': 'BASELINE configs[1] at full shape: %d nt, default pipeline, every round vs oracle/dp_oracle.c (OpenMP, %d threads)'
                              % (args.nt, threads), 'all_rounds_equal': all_equal, 'rounds': rounds, 'final': final,
                  'segments': int(len(scores)), 'score': float(total), 'host_tables_sha1': None, 'seconds': round(time.time() - t0, 1)}))
