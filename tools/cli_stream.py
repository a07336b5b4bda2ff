"""Config 5 through the text boundary: many short contigs as bedgraph text -> split_bedgraph_stream -> segment text.
python tools/cli_stream.py [n_contigs]"""
import io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth, process_bedgraph as pb
from pasio_b200.splitters import configure_splitter
n_contigs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
lens = synth.transcript_lengths(n_contigs)
t0 = time.time()
lines = []
for j, n in enumerate(lens):
    lines.extend(synth.to_bedgraph_lines('tx%d' % j, synth.dnase_like(int(n), seed=5000 + j, hotspot_share=0.3)))
text = ''.join(lines)
print('generated %d contigs, %.3g nt, %.1f MB of bedgraph text in %.1f s' % (n_contigs, lens.sum(), len(text) / 1e6, time.time() - t0), flush=True)
splitter = configure_splitter()
data = text.encode()
del text, lines
for rep in range(2):
    out = io.StringIO()
    t0 = time.perf_counter()
    pb.split_bedgraph_stream(io.TextIOWrapper(io.BytesIO(data)), out, splitter)       # like a file opened in text mode
    dt = time.perf_counter() - t0
    print('run %d: %.2f s -> %.3g nt/s, %.0f contigs/s, %d output lines' % (rep, dt, lens.sum() / dt, n_contigs / dt, out.getvalue().count('\n')), flush=True)
