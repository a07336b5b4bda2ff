"""BASELINE config 5 at its stated size through the text boundary: 100 000 transcript-like contigs (1-50 kb, 2.55e9 nt) as
run-length bedgraph text on STDIN of `python -m pasio_b200 -`, segment text on stdout.  Reports wall time, nt/s, contigs/s,
the child's peak resident memory (bounded: the input is streamed in 64 MB pieces) and -- in-process -- when the first
output was written relative to the end of the input.

    python tools/config5_stdin.py [--contigs 100000] > gpurun_out/config5_stdin.json
"""
import argparse
import io
import json
import multiprocessing as mp
import os
import resource
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pasio_b200 import synth                                    # noqa: E402


def _text_of(job):
    lo, hi, lens = job
    out = []
    for j in range(lo, hi):
        out.extend(synth.to_bedgraph_lines('tx%06d' % j, synth.dnase_like(int(lens[j]), seed=5000 + j, hotspot_share=0.3)))
    return ''.join(out).encode()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--contigs', type=int, default=100000)
    args = ap.parse_args()
    lens = synth.transcript_lengths(args.contigs)
    t0 = time.time()
    step = 250
    jobs = [(lo, min(lo + step, args.contigs), lens) for lo in range(0, args.contigs, step)]
    path = os.path.join(tempfile.mkdtemp(dir='/dev/shm' if os.path.isdir('/dev/shm') else None), 'config5.bedgraph')
    with mp.get_context('fork').Pool(os.cpu_count() or 1) as pool, open(path, 'wb') as f:
        for piece in pool.imap(_text_of, jobs):
            f.write(piece)
    text_bytes = os.path.getsize(path)
    gen_s = time.time() - t0
    out_path = path + '.out'
    # (1) the CLI as a child process, text piped on stdin
    runs = []
    for rep in range(2):
        t0 = time.perf_counter()
        with open(path, 'rb') as src, open(out_path, 'wb') as dst:
            child = subprocess.Popen([sys.executable, '-m', 'pasio_b200', '-', '-o', '-'], stdin=subprocess.PIPE, stdout=dst, cwd=ROOT)
            while True:
                piece = src.read(1 << 24)
                if not piece:
                    break
                child.stdin.write(piece)
            child.stdin.close()
            child.wait()
        dt = time.perf_counter() - t0
        assert child.returncode == 0
        runs.append(dt)
    rss_mb = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1024.0
    n_lines = sum(1 for _ in open(out_path, 'rb'))
    # (2) in-process: when does the first output appear?
    import pasio_b200
    from pasio_b200 import process_bedgraph as pb
    marks = {}

    class Reader(io.RawIOBase):
        def __init__(self, f):
            self.f = f

        def readable(self):
            return True

        def readinto(self, b):
            data = self.f.read(len(b))
            if not data and 'eof' not in marks:
                marks['eof'] = time.perf_counter()
            b[:len(data)] = data
            return len(data)

    class Writer(io.RawIOBase):
        def writable(self):
            return True

        def write(self, b):
            marks.setdefault('first_write', time.perf_counter())
            return len(b)

    t0 = time.perf_counter()
    with open(path, 'rb') as f:
        pb.split_bedgraph_stream(io.TextIOWrapper(io.BufferedReader(Reader(f), 1 << 20)), io.TextIOWrapper(io.BufferedWriter(Writer())),
                                 pasio_b200.configure_splitter())
    inproc = time.perf_counter() - t0
    nt = int(lens.sum())
    best = min(runs)
    print(json.dumps({
        'workload': 'BASELINE configs[4]: %d contigs of 1-50 kb (%d nt) as %.2f GB of run-length bedgraph text piped on stdin of '
                    '`python -m pasio_b200 - -o -`; %d output lines' % (args.contigs, nt, text_bytes / 1e9, n_lines),
        'cli_seconds': runs, 'nt_per_s': nt / best, 'contigs_per_s': args.contigs / best, 'text_MB_per_s': text_bytes / 1e6 / best,
        'child_peak_rss_MB': rss_mb, 'input_MB': text_bytes / 1e6,
        'in_process_seconds': inproc, 'first_output_after_s': marks.get('first_write', t0) - t0,
        'input_eof_after_s': marks.get('eof', t0) - t0, 'host_text_generation_s': gen_s}))
    os.unlink(path)
    os.unlink(out_path)


if __name__ == '__main__':
    main()
