"""Randomised differential test of the rounds path (incl. the branch-and-bound pruning) against the C oracle.
python tests/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth, _native
from pasio_b200.log_marginal_likelyhood import ScorerFactory
from oracle import c_oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rs = np.random.RandomState(seed)
eng = _native.engine()
factories = {}
t_end = time.time() + budget
cases = skipped_cells = total_cells = 0
while time.time() < t_end:
    kind = rs.randint(0, 6)
    n = int(rs.choice([3000, 20000, 60000, 150000]))
    if kind == 0:
        counts = synth.dnase_like(n, int(rs.randint(1 << 30)), hotspot_share=float(rs.uniform(0.05, 0.9)))
    elif kind == 1:
        counts = synth.piecewise_poisson(n, int(rs.randint(1 << 30)))
    elif kind == 2:
        counts = (rs.random_sample(n) < rs.uniform(0.01, 0.5)).astype(np.int64)              # 0/1: many exact ties
    elif kind == 3:
        counts = (synth.piecewise_poisson(n, int(rs.randint(1 << 30))) * int(rs.randint(1, 2000))).astype(np.int64)
    elif kind == 4:
        counts = np.repeat(rs.poisson(rs.uniform(0.2, 30), n // 50 + 1), 50)[:n].astype(np.int64)   # long constant runs
    else:
        counts = rs.randint(0, 3, n).astype(np.int64)
    alpha = float(rs.choice([1.0, 1.0, 2.0, 0.5, 0.01, 3.7, 25.0]))
    beta = float(rs.choice([1.0, 1.0, 0.1, 2.5, 10.0]))
    wsize = int(rs.choice([150, 159, 160, 300, 511, 512, 700, 2500, 2500, 5000]))   # 159/160, 511/512: size-class edges of the window kernels
    wshift = int(rs.choice([wsize // 2, wsize // 2, wsize // 3 + 1, wsize]))
    constraint = str(rs.choice(['constants', 'constants', 'none', 'zeros']))
    if constraint != 'constants' and n > 60000:
        n = 60000
        counts = counts[:n]
    key = (alpha, beta)
    if key not in factories:
        factories[key] = ScorerFactory(alpha, beta)
    fo = c_oracle.FlatOracle(counts, alpha, beta)
    eng.use_scorer(factories[key])
    eng.load(counts)
    if rs.random_sample() < 0.5:
        # explicit first list: a random subset of the positions -- noisy candidates through the pruned (pipelined) path
        inner = np.flatnonzero(rs.random_sample(n - 1) < rs.uniform(0.03, 0.6)) + 1
        cands = np.concatenate([[0], inner, [n]]).astype(np.int64)
        eng.set_candidates(cands)
    else:
        eng.set_candidates(None)
        cands = np.arange(n + 1, dtype=np.int64)
    for r in range(4):
        eng.round(wsize, wshift, constraint)
        got = eng.candidates()
        want, o_cells = fo.round(cands, wsize, wshift, constraint)
        c, sk = eng.round_stats()
        total_cells += c
        skipped_cells += sk
        if not np.array_equal(got, want) or c != o_cells:
            print('MISMATCH kind=%d n=%d alpha=%g beta=%g w=%d/%d %s round=%d seed=%d' % (kind, n, alpha, beta, wsize, wshift, constraint, r, seed))
            sys.exit(1)
        if len(want) == len(cands):
            break
        cands = want
    cases += 1
print('fuzz ok: %d cases, %.4g cells, %.1f%% skipped by the bound' % (cases, total_cells, 100.0 * skipped_cells / max(1, total_cells)))
