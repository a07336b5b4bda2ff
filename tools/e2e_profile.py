"""Where does the host-side time of one end-to-end contig go?  python tools/e2e_profile.py --nt N"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasio_b200 import synth, _native
from pasio_b200.splitters import configure_splitter, _fusion
from pasio_b200.segmentation import _run_device_pipeline

ap = argparse.ArgumentParser(); ap.add_argument('--nt', type=int, default=248956422); args = ap.parse_args()
n = args.nt
host = torch.empty(n, dtype=torch.int64, pin_memory=True); counts = host.numpy(); counts[:] = synth.dnase_like(n, seed=1000)
eng = _native.engine(); plan = _fusion.pipeline_plan(configure_splitter())
for rep in range(3):
    eng.invalidate(); T = [time.perf_counter()]; names = []
    def mark(name):
        T.append(time.perf_counter()); names.append(name)
    eng.use_scorer(plan['factory']); mark('use_scorer')
    eng.load(counts); mark('load')
    eng.set_candidates(None); mark('set_cands')
    _run_device_pipeline(eng, plan); mark('rounds')
    splits = eng.candidates(); mark('candidates d2h')
    scores, _, means, logfac = eng.segment_scores(scores=True, means=True, logfac=True); mark('segment_scores')
    score = np.sum(scores); lmm = scores - np.diff(logfac); mark('numpy tail')
    print('rep', rep, ' '.join('%s=%.1fms' % (nm, (b - a) * 1e3) for nm, a, b in zip(names, T[:-1], T[1:])), 'total=%.1fms' % ((T[-1] - T[0]) * 1e3), flush=True)
