"""Wall-clock breakdown of one device-resident bench step (host gaps vs kernel time)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pasio_b200 import synth, _native
from pasio_b200.splitters import configure_splitter, _fusion
from pasio_b200.segmentation import _run_device_pipeline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 248956422
host = torch.empty(n, dtype=torch.int64, pin_memory=True)
host.numpy()[:] = synth.dnase_like(n, seed=1000)
resident = host.cuda(); torch.cuda.synchronize()
eng = _native.engine()
plan = _fusion.pipeline_plan(configure_splitter())
for rep in range(4):
    eng.timing_reset(True)
    t = [time.perf_counter()]
    eng.use_scorer(plan['factory']); eng.load_device(resident.data_ptr(), n, owner=resident); t.append(time.perf_counter())
    eng.set_candidates(None); t.append(time.perf_counter())
    _run_device_pipeline(eng, plan); t.append(time.perf_counter())
    scores, _, means, _ = eng.segment_scores(scores=True, means=True); t.append(time.perf_counter())
    s = float(np.sum(scores)); t.append(time.perf_counter())
    tm = eng.timing()
    names = ['load_device', 'set_candidates', 'rounds', 'segment_scores', 'np.sum']
    print('rep %d wall ms: ' % rep + ', '.join('%s %.2f' % (nm, (b - a) * 1e3) for nm, a, b in zip(names, t[:-1], t[1:])) + ' | total %.2f' % ((t[-1] - t[0]) * 1e3))
    print('      device ms: ' + ', '.join('%s %.2f (%d)' % (k, v[0], v[1]) for k, v in tm.items()))
