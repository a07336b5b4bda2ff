"""The other BASELINE.json configs (bench.py times configs[1]); one JSON line per run.

  python tools/workloads.py exact1                      # config 1: exact DP, n=100 000, all positions
  python tools/workloads.py exact3                      # config 3: exact DP, N=200 000 candidates over 2 Mb
  python tools/workloads.py genome [--scale 1.0]        # config 4: hg38 contig-size profile, LPT over the ranks
  python tools/workloads.py transcripts [--contigs N]   # config 5: short contigs batched per launch
Under torchrun (WORLD_SIZE > 1) `genome` shards the contigs by LPT (pasio_b200.sharding) and reports the
max-over-ranks time; the others are single-GPU."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasio_b200 import synth, _native, sharding                 # noqa: E402
from pasio_b200.splitters import configure_splitter, _fusion   # noqa: E402
from pasio_b200.segmentation import segment_on_device           # noqa: E402
from pasio_b200.log_marginal_likelyhood import ScorerFactory    # noqa: E402


def exact(args, which):
    eng = _native.engine()
    eng.use_scorer(ScorerFactory(1.0, 1.0))
    if which == 1:
        counts = synth.piecewise_poisson(100000, 0)
        cands = None
        N = len(counts) + 1
    else:
        counts = synth.piecewise_poisson(2000000, 1)
        cands = synth.random_candidates(len(counts), 200000, 1)
        N = len(cands)
    eng.load(counts)
    eng.set_tuning('exact_prune', args.prune)
    eng.set_tuning('exact_nblock', args.nblock)
    eng.set_tuning('exact_lag', args.lag)
    eng.set_tuning('exact_ring', args.ring)
    times = []
    for rep in range(args.reps + 1):
        eng.set_candidates(cands)
        eng._cands_obj = None
        eng.timing_reset(True)
        t0 = time.perf_counter()
        score, splits = eng.square_split()
        wall = time.perf_counter() - t0
        times.append((wall, eng.timing()['exact_dp'][0]))
    wall, kern = min(times[1:] if len(times) > 1 else times)
    cells = N * (N - 1) // 2
    _, skipped = eng.round_stats()
    peak = 148 * 64 * 1.965e9
    print(json.dumps({'workload': 'config%d exact SquareSplitter' % which, 'N': N, 'cells': cells,
                      'prune': args.prune, 'lag': args.lag, 'cells_skipped': skipped,
                      'cells_evaluated': cells - skipped, 'evaluated_frac': (cells - skipped) / cells,
                      'fp64_roofline_frac_algorithmic': cells / (kern * 1e-3) * 4 / peak,
                      'fp64_roofline_frac_evaluated': (cells - skipped) / (kern * 1e-3) * 4 / peak,
                      'ns_per_row': kern * 1e6 / N,
                      'wall_ms': wall * 1e3, 'kernel_ms': kern, 'cells_per_s_wall': cells / wall,
                      'cells_per_s_kernel': cells / (kern * 1e-3), 'splits': len(splits), 'score': float(score)}))


def genome(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group("gloo")      # control plane and the host-side gather only: contigs are partitioned, never exchanged
    sizes = synth.genome_profile(scale=args.scale)
    costs = [sharding.contig_cost(n) for _, n in sizes]
    mine = sharding.shard_indices(costs, rank, world)
    plan = _fusion.pipeline_plan(configure_splitter())
    eng = _native.engine()
    contigs = {i: synth.dnase_like(sizes[i][1], seed=i) for i in mine}      # host generation, untimed
    if dist is not None:
        dist.barrier()
    # scaffolds and other short contigs share one batched load (one launch per round for all of them)
    small = [i for i in mine if sizes[i][1] < (16 << 20)]
    big = [i for i in mine if sizes[i][1] >= (16 << 20)]
    t0 = time.perf_counter()
    nseg = 0
    for i in big:
        eng.invalidate()
        score, splits, means, lmm, _ = segment_on_device(contigs[i], plan, want_lmm=False)
        nseg += len(splits) - 1
    if small:
        offsets = np.concatenate([[0], np.cumsum([sizes[i][1] for i in small])]).astype(np.int64)
        batch = np.concatenate([contigs[i] for i in small])
        eng.use_scorer(plan['factory'])
        eng.load(batch, offsets=offsets)
        eng.set_candidates(None)
        _, final, _ = eng.rounds(2500, 1250, 'constants')
        scores, _, means, _ = eng.segment_scores(scores=True, means=True)
        splits = eng.candidates()
        nseg += final - 1
    dt = time.perf_counter() - t0
    total_nt = sum(n for _, n in sizes)
    if dist is not None:
        import torch
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(nseg)], dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dt, nseg = float(t.item()), int(s.item())
    if rank == 0:
        print(json.dumps({'workload': 'config4 genome profile (scale %g): %d contigs, default pipeline, host counts -> segments'
                                      % (args.scale, len(sizes)), 'n_gpus': world, 'nt': total_nt, 'seconds': dt,
                          'nt_per_s': total_nt / dt, 'segments': nseg,
                          'lpt_max_over_mean': max(sum(costs[i] for i in sharding.shard_indices(costs, r, world))
                                                   for r in range(world)) / (sum(costs) / world)}))
    if dist is not None:
        dist.destroy_process_group()


def transcripts(args):
    lens = synth.transcript_lengths(args.contigs)
    eng = _native.engine()
    f = ScorerFactory(1.0, 1.0)
    eng.use_scorer(f)
    batch_nt = 1 << 30
    total_nt, total_t, total_seg, batches = 0, 0.0, 0, 0
    k = 0
    while k < len(lens):
        k1, acc = k, 0
        while k1 < len(lens) and acc + lens[k1] <= batch_nt:
            acc += int(lens[k1])
            k1 += 1
        offsets = np.concatenate([[0], np.cumsum(lens[k:k1])]).astype(np.int64)
        counts = np.concatenate([synth.dnase_like(int(n), seed=5000 + k + j, hotspot_share=0.3)
                                 for j, n in enumerate(lens[k:k1])])
        t0 = time.perf_counter()
        eng.load(counts, offsets=offsets)
        eng.set_candidates(None)
        sizes, final, cells = eng.rounds(2500, 1250, 'constants')
        scores, _, means, _ = eng.segment_scores(scores=True, means=True)
        splits = eng.candidates()
        dt = time.perf_counter() - t0
        total_nt += acc
        total_t += dt
        total_seg += final - 1
        batches += 1
        k = k1
    print(json.dumps({'workload': 'config5 transcript-like stream: %d contigs of 1-50 kb, batched per launch' % len(lens),
                      'batches': batches, 'nt': total_nt, 'seconds': total_t, 'nt_per_s': total_nt / total_t,
                      'contigs_per_s': len(lens) / total_t, 'segments': total_seg, 'rounds_last_batch': len(sizes)}))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['exact1', 'exact3', 'genome', 'transcripts'])
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--contigs', type=int, default=100000)
    ap.add_argument('--prune', type=int, default=1, help='exact DP: 1 = bounded far columns (default), 0 = every cell')
    ap.add_argument('--ring', type=int, default=1, help='exact DP: 1 = self scores in the ring layout of very long lists')
    ap.add_argument('--lag', type=int, default=5, help='exact DP: far columns start this many blocks behind (3 .. 5)')
    ap.add_argument('--nblock', type=int, default=3, help='exact DP: column blocks in front of the far columns evaluated by worker CTAs (0 .. lag - 2)')
    a = ap.parse_args()
    if a.what == 'exact1':
        exact(a, 1)
    elif a.what == 'exact3':
        exact(a, 3)
    elif a.what == 'genome':
        genome(a)
    else:
        transcripts(a)
