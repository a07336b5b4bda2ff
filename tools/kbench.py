"""Kernel-level experiment harness (not the bench): times the window-DP rounds on one synthetic
contig with the library's own CUDA-event hooks.  Usage: python tools/kbench.py --nt 100000000"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth, _native                      # noqa: E402
from pasio_b200.log_marginal_likelyhood import ScorerFactory  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--nt', type=int, default=100000000)
ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--exact', type=int, default=0, help='also time the exact DP with this many nt (all positions)')
ap.add_argument('--config3', type=int, default=0, help='time the exact DP with this many random candidates over 10x as many nt')
args = ap.parse_args()

eng = _native.engine()
f = ScorerFactory(1.0, 1.0)
eng.use_scorer(f)
counts = synth.dnase_like(args.nt, seed=1000)
eng.load(counts)
for rep in range(args.reps):
    eng.set_candidates(None)
    eng.timing_reset(True)
    t0 = time.perf_counter()
    sizes, final, cells = eng.rounds(2500, 1250, 'constants')
    wall = time.perf_counter() - t0
    t = eng.timing()
    print('rep %d: rounds=%d cells=%.4g wall=%.1f ms window_dp=%.1f ms (%d launches) -> %.4g cells/s kernel, compact %.2f ms'
          % (rep, len(sizes), cells, wall * 1e3, t['window_dp'][0], t['window_dp'][1], cells / (t['window_dp'][0] * 1e-3),
             t['compact'][0]), flush=True)
if args.exact:
    c = synth.piecewise_poisson(args.exact, 0)
    eng.load(c)
    for rep in range(2):
        eng.set_candidates(None)
        eng.timing_reset(True)
        t0 = time.perf_counter()
        score, splits = eng.square_split()
        wall = time.perf_counter() - t0
        t = eng.timing()
        N = args.exact + 1
        print('exact N=%d: wall %.1f ms, exact_dp %.1f ms (%d launches) -> %.4g cells/s, splits %d score %.6f'
              % (N, wall * 1e3, t['exact_dp'][0], t['exact_dp'][1], N * (N - 1) / 2 / (t['exact_dp'][0] * 1e-3), len(splits), score))

if args.config3:
    N = args.config3
    c = synth.piecewise_poisson(10 * N, 1)
    cands = synth.random_candidates(len(c), N, 1)
    eng.load(c)
    for rep in range(2):
        eng.set_candidates(cands)
        eng._cands_obj = None
        eng.timing_reset(True)
        t0 = time.perf_counter()
        score, splits = eng.square_split()
        wall = time.perf_counter() - t0
        t = eng.timing()
        print('config3 N=%d: wall %.1f ms, exact_dp %.1f ms -> %.4g cells/s, splits %d score %.6f'
              % (N, wall * 1e3, t['exact_dp'][0], N * (N - 1) / 2 / (t['exact_dp'][0] * 1e-3), len(splits), score))
