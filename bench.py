#!/usr/bin/env python
"""bench.py -- throughput of the Pasio segmentation hot path on B200.

A "step" is one pass of the default pasio pipeline (NotConstantReducer + RoundReducer over
SlidingWindowReducer(2500, 1250) + SquareSplitter, then NopSplitter scoring) over one synthetic
hg38-chr1-sized DNase-like contig (BASELINE.json configs[1]).  With N GPUs every rank segments
its own chr1-sized contig (contigs are independent: no data-path collective; weak scaling).

  value : whole-job nt/s with the counts already resident in HBM when the timed region starts
  e2e   : the same through the public API (pasio_b200.segmentation.segment_on_device) from a
          PINNED HOST int64 buffer: H2D of the counts and D2H of splits / scores / means / logfac
          inside the timed region
  roofline : the dominant kernel (batched window DP), 4 FP64 ops per (i,j) cell against
          the FP64-pipe instruction rate
  cpu_baseline : oracle/ (numpy port of the reference) on a bounded prefix of the same contig

`--impl reference` times the reference's CPU algorithm (oracle port; the reference is pure Python
and /root/reference does not exist on the GPU box) on all host cores, same metric and config.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHR1 = 248956422
WINDOW_SIZE, WINDOW_SHIFT, CONSTRAINT = 2500, 1250, 'constants'
FP64_OPS_PER_CELL = 4            # DMUL s*Lg, DADD G-.., DADD +P_i, compare


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return json.load(open(path)), 'measured'
    return {'hbm_gbs': 6650.0, 'sm_max_mhz': 1965.0}, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        fd, self.path = tempfile.mkstemp(suffix='.csv')
        os.close(fd)
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend='nccl', device_id=torch.device('cuda', local))
    return rank, world, local, dist


def max_over_ranks(x, dist, device):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, dist, device):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(dist, torch):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    from pasio_b200 import synth, _native
    from pasio_b200.splitters import configure_splitter, _fusion
    from pasio_b200.segmentation import segment_on_device, _run_device_pipeline

    rank, world, local, dist = dist_setup(args.gpus)
    os.environ.setdefault('PASIO_B200_DEVICE', str(local))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    n = args.nt
    peaks, peak_kind = measured_peaks()

    # synthetic input, generated straight into pinned host memory (one chr1-sized contig per rank)
    host = torch.empty(n, dtype=torch.int64, pin_memory=True)
    counts = host.numpy()
    counts[:] = synth.dnase_like(n, seed=1000 + rank)
    resident = host.to(device, non_blocking=False)          # the HBM-resident copy for `value`
    torch.cuda.synchronize()

    eng = _native.engine()
    splitter = configure_splitter(window_size=WINDOW_SIZE, window_shift=WINDOW_SHIFT, split_constraints=CONSTRAINT)
    plan = _fusion.pipeline_plan(splitter)
    assert plan is not None
    stream = torch.cuda.ExternalStream(eng.stream_handle(), device=device)

    def step_resident():
        eng.use_scorer(plan['factory'])
        eng.load_device(resident.data_ptr(), n, owner=resident)
        eng.set_candidates(None)
        _run_device_pipeline(eng, plan)
        scores, _, means, _ = eng.segment_scores(scores=True, means=True)
        return eng.candidate_count(), float(eng.segment_scores_sum())

    def step_e2e():
        eng.invalidate()                                     # force the H2D copy every step
        score, splits, means, lmm, _ = segment_on_device(counts, plan, want_lmm=True)
        return len(splits), float(score)

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(args.warmup):
        m_final, score = step_resident()
    fp64_peak_measured = eng.fp64_peak() if rank == 0 else 0.0
    eng.timing_reset(True)
    sampler = ClockSampler(local)
    barrier(dist, torch)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m_final, score = step_resident()
    ev1.record(stream)
    ev1.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    barrier(dist, torch)
    dev_ms = ev0.elapsed_time(ev1)
    timing = eng.timing()
    eng.timing_reset(False)
    step_ms = max_over_ranks(dev_ms / args.steps, dist, device)
    total_nt = sum_over_ranks(float(n), dist, device)
    value = total_nt / (step_ms * 1e-3)

    # cells of one step (same every step): one extra instrumented pass
    eng.use_scorer(plan['factory'])
    eng.load_device(resident.data_ptr(), n, owner=resident)
    eng.set_candidates(None)
    sizes, final, cells = eng.rounds(WINDOW_SIZE, WINDOW_SHIFT, CONSTRAINT)

    # ---- end to end from pinned host memory ----------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier(dist, torch)
    eng.timing_reset(True)
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    e2e_each = []
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        m_e2e, score_e2e = step_e2e()
        e2e_each.append((time.perf_counter() - t1) * 1e3)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_timing = eng.timing()
    eng.timing_reset(False)
    e2e_s = max_over_ranks(e2e_s, dist, device)
    e2e_value = total_nt / e2e_s
    assert m_e2e == m_final and abs(score_e2e - score) <= 1e-9 * abs(score)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------
    wd_ms, wd_launches = timing['window_dp']
    scan_ms, scan_launches = timing['scan']
    fp64_peak_nominal = 148 * 64 * peaks.get('sm_max_mhz', 1965.0) * 1e6
    cells_per_launch = cells / max(1, len(sizes))
    # one timed span per round: the CTA-per-window kernel and the two warp-per-window kernels of a round overlap
    # (side stream), so the per-round span is the unit; wd_launches counts the kernels themselves
    wd_rounds = max(1, len(sizes) * args.steps)
    wd_avg_s = (wd_ms / wd_rounds) * 1e-3
    achieved = cells_per_launch * FP64_OPS_PER_CELL / wd_avg_s
    traffic = None
    prof = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(prof):
        traffic = json.load(open(prof)).get('window_dp_dram_bytes_per_launch')
    roofline = {
        'kernel': 'window DP of one round: window_dp_kernel (CTA per window) + small_window_dp_kernel x2 (warp per window, '
                  'side stream), timed as one span per round',
        'bound': 'fp64', 'achieved': achieved / 1e12, 'peak': fp64_peak_nominal / 1e12, 'unit': 'TFLOP/s',
        'frac': achieved / fp64_peak_nominal, 'traffic': traffic,
        'peak_source': '148 SM x 64 FP64 lanes x sm_max_mhz (%s MEASURED_PEAKS.json has no FP64 entry); '
                       'DFMA micro-benchmark on this box: %.3g instr/s' % (peak_kind, fp64_peak_measured),
        'algorithmic_ops_per_cell': FP64_OPS_PER_CELL, 'cells_per_launch': cells_per_launch,
        'avg_launch_ms': wd_avg_s * 1e3, 'launches': wd_rounds, 'kernels_launched': wd_launches,
        'share_of_step': wd_ms / (dev_ms if dev_ms > 0 else 1.0),
        'scan_kernel': {'bound': 'hbm', 'achieved': 16.0 * n / ((scan_ms / max(1, scan_launches)) * 1e-3) / 1e9,
                        'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                        'frac': 16.0 * n / ((scan_ms / max(1, scan_launches)) * 1e-3) / 1e9 / peaks['hbm_gbs'],
                        'peak_source': peak_kind},
    }

    # ---- CPU baseline: the oracle port on a bounded prefix of the same contig ------------------
    cpu = cpu_baseline_port(counts, args.cpu_sample_nt)

    gpu_launches = int(sum(timing[k][1] for k in ['scan', 'window_dp', 'compact', 'exact_dp', 'score']))
    line = {
        'metric': 'whole-contig segmentation throughput, default pasio pipeline (nt/s); DP cell updates/s in dp_cells_per_s',
        'value': value, 'unit': 'nt/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1]: default pasio pipeline (constants + rounds over sliding window '
                               '2500/1250 + SquareSplitter, alpha=beta=1) on one synthetic chr1-sized (%d nt) DNase-like '
                               'contig per GPU' % n,
                   'nt_per_gpu': n, 'window_size': WINDOW_SIZE, 'window_shift': WINDOW_SHIFT,
                   'split_constraints': CONSTRAINT, 'rounds': len(sizes), 'candidates_per_round': sizes,
                   'segments': final - 1, 'l2_policy': 'inputs larger than L2 (2 GB counts + 2 GB prefix sums per step)',
                   'parallelism': 'contigs sharded over %d GPU(s), LPT, no collective' % world},
        'dp_cells_per_step': cells, 'dp_cells_per_s': cells * world / (step_ms * 1e-3),
        'window_dp_cells_per_s_kernel_only': cells / (wd_ms / args.steps * 1e-3),
        'e2e': {'value': e2e_value, 'unit': 'nt/s', 'h2d_bytes_per_step': int(n * 8),
                'd2h_bytes_per_step': int(final * 8 * 4), 'ms_per_step': e2e_s * 1e3,
                'api': 'pasio_b200.segmentation.segment_on_device(counts_pinned_host, plan)',
                'device_ms_per_step': {k: e2e_timing[k][0] / e2e_steps for k in e2e_timing},
                'ms_each_step': e2e_each},
        'gpu_launches': gpu_launches,
        'kernel_ms_per_step': {k: timing[k][0] / args.steps for k in timing},
        'wall_ms_per_step': wall / args.steps * 1e3,
        'clocks': {'sm_mhz': clocks['sm_mhz'], 'sm_max_mhz': clocks['sm_max_mhz'], 'reasons': clocks['reasons'],
                   'samples': clocks['samples']},
        'roofline': roofline, 'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_baseline_port(counts, sample_nt):
    """oracle/pasio_oracle.py (numpy restatement, one core) on the first sample_nt nt."""
    from oracle import pasio_oracle as po
    sample = np.ascontiguousarray(counts[:sample_nt])
    tables = po.Tables(1, 1.0)
    t0 = time.perf_counter()
    out = po.default_pipeline(sample, tables, WINDOW_SIZE, WINDOW_SHIFT, CONSTRAINT)
    dt = time.perf_counter() - t0
    return {'value': len(sample) / dt, 'unit': 'nt/s', 'cores': 1, 'kind': 'port',
            'sample': 'first %d nt of the rank-0 workload contig through oracle/pasio_oracle.default_pipeline '
                      '(numpy port of the reference, %d segments, %.1f s)' % (len(sample), len(out['splits']) - 1, dt)}


_REF_COUNTS = None      # set in the parent before the pool forks


def _ref_worker(job):
    lo, hi = job
    from oracle import pasio_oracle as po
    counts = _REF_COUNTS[lo:hi]
    tables = po.Tables(1, 1.0)
    t0 = time.perf_counter()
    out = po.default_pipeline(np.ascontiguousarray(counts), tables, WINDOW_SIZE, WINDOW_SHIFT, CONSTRAINT)
    return hi - lo, time.perf_counter() - t0, len(out['splits']) - 1


def run_reference(args):
    """The reference's CPU algorithm (oracle port) on every host core: one slice of the workload contig
    per core per step, the process-per-contig model of tests/pasio_parallel_wrapper.py."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    slice_nt = args.ref_slice_nt
    head = min(args.nt, cores * slice_nt * (args.steps + args.warmup))
    jobs_per_step = cores
    global _REF_COUNTS
    from pasio_b200 import synth
    _REF_COUNTS = synth.dnase_like(head, seed=1000)
    pool = mp.get_context('fork').Pool(cores)
    step_times, nt_per_step = [], jobs_per_step * slice_nt
    k = 0
    for step in range(args.warmup + args.steps):
        jobs = []
        for c in range(jobs_per_step):
            lo = (k * slice_nt) % max(slice_nt, head - slice_nt + 1)
            jobs.append((lo, lo + slice_nt))
            k += 1
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, jobs)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            step_times.append(dt)
    pool.close()
    ms = float(np.mean(step_times)) * 1e3
    value = nt_per_step / (ms * 1e-3)
    sample = ('%d slices of %d nt of the workload contig per step, one per host core, each through '
              'oracle/pasio_oracle.default_pipeline (numpy port of the reference; the reference itself is pure '
              'Python under /root/reference, which does not exist on the GPU box)' % (jobs_per_step, slice_nt))
    line = {
        'impl': 'reference',
        'metric': 'whole-contig segmentation throughput, default pasio pipeline (nt/s); DP cell updates/s in dp_cells_per_s',
        'value': value, 'unit': 'nt/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1]: default pasio pipeline on synthetic chr1-sized DNase-like contig '
                               '(bounded sample per step)', 'window_size': WINDOW_SIZE, 'window_shift': WINDOW_SHIFT,
                   'split_constraints': CONSTRAINT},
        'cpu_baseline': {'value': value, 'unit': 'nt/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'nt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nt', type=int, default=CHR1, help='contig length per GPU (default: hg38 chr1)')
    ap.add_argument('--cpu-sample-nt', type=int, default=2000000)
    ap.add_argument('--ref-slice-nt', type=int, default=500000)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
