"""End-to-end CLI timing: bedgraph text file -> python -m pasio_b200 -> segment text file."""
import os, subprocess, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 100000000
d = tempfile.mkdtemp()
src, dst = os.path.join(d, 'in.bedgraph'), os.path.join(d, 'out.bedgraph')
c = synth.dnase_like(nt, seed=1000)
t0 = time.perf_counter()
with open(src, 'w') as f:
    f.writelines(synth.to_bedgraph_lines('chr1', c))
print('wrote %s: %.1f MB, %d lines in %.1f s' % (src, os.path.getsize(src) / 1e6, sum(1 for _ in open(src)), time.perf_counter() - t0), flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    subprocess.check_call([sys.executable, '-m', 'pasio_b200', src, '-o', dst], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dt = time.perf_counter() - t0
    print('CLI run %d: %.2f s for %d nt -> %.3g nt/s, %d segments' % (rep, dt, nt, nt / dt, sum(1 for _ in open(dst))), flush=True)
# in-process (no interpreter / CUDA start-up)
import pasio_b200
s = pasio_b200.configure_splitter()
for rep in range(2):
    t0 = time.perf_counter()
    pasio_b200.split_bedgraph(src, dst, s)
    dt = time.perf_counter() - t0
    print('in-process split_bedgraph %d: %.2f s -> %.3g nt/s' % (rep, dt, nt / dt), flush=True)
