"""Multi-GPU execution: contigs are independent, so they are partitioned, not exchanged.

One process per GPU (torchrun sets RANK / LOCAL_RANK / WORLD_SIZE); contigs are assigned by
longest-processing-time-first, the ordering heuristic of the reference's per-chromosome script
generator (/root/reference/tests/pasio_parallel_wrapper.py:68-74, "sort by length desc"); every
rank segments its contigs on its own GPU; the per-contig results are gathered ON THE HOST in
input order: through shard files when a directory is given (the default of the product path,
pasio_b200/device_pool.py does the same for `split_bedgraph(..., devices=N)`), else through a
gloo (CPU) process group.  Nothing of the data path goes over NCCL.
"""
import heapq

import numpy as np


def lpt_assign(costs, n_ranks):
    """Longest-processing-time-first: returns rank_of[i] for every contig i (deterministic)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0.0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    rank_of = np.zeros(len(costs), dtype=np.int64)
    for i in order:
        load, r = heapq.heappop(heap)
        rank_of[i] = r
        heapq.heappush(heap, (load + costs[i], r))
    return rank_of


def contig_cost(length):
    """Cost model for LPT: a contig's work grows with its length (windows per round ~ length)."""
    return float(length)


def shard_indices(costs, rank, world_size):
    rank_of = lpt_assign(costs, world_size)
    return [i for i in range(len(costs)) if rank_of[i] == rank]


_GLOO = {}


def _host_group(dist):
    """a gloo (CPU) group over all ranks: host objects are gathered on the host, whatever the default backend is"""
    if dist.get_backend() == 'gloo':
        return None
    if 'group' not in _GLOO:
        _GLOO['group'] = dist.new_group(backend='gloo')
    return _GLOO['group']


def gather_in_order(local_results, costs, rank, world_size, dist=None, shard_dir=None):
    """local_results: {contig index: result} of this rank.  Returns the full list in input order
    on rank 0 (None elsewhere).  dist: torch.distributed (initialised) or None for world_size 1.
    shard_dir: a directory every rank can write to -- results travel as one pickle file per rank
    (rank 0 reads them after a barrier); without it they go through a gloo group."""
    if world_size == 1 or dist is None:
        return [local_results[i] for i in range(len(costs))]
    if shard_dir is not None:
        import os
        import pickle
        with open(os.path.join(shard_dir, 'shard_%05d.pkl' % rank), 'wb') as f:
            pickle.dump(local_results, f, protocol=pickle.HIGHEST_PROTOCOL)
        dist.barrier(group=_host_group(dist))
        if rank != 0:
            return None
        merged = {}
        for r in range(world_size):
            with open(os.path.join(shard_dir, 'shard_%05d.pkl' % r), 'rb') as f:
                merged.update(pickle.load(f))
        return [merged[i] for i in range(len(costs))]
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(local_results, gathered, dst=0, group=_host_group(dist))
    if rank != 0:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(costs))]


def segment_contigs(contigs, segment_fn, rank=0, world_size=1, dist=None, shard_dir=None):
    """contigs: list of (name, counts, start).  segment_fn(name, counts, start) -> result (e.g. TSV text).
    Every rank runs its LPT share; rank 0 receives all results in input order."""
    costs = [contig_cost(len(c[1])) for c in contigs]
    mine = shard_indices(costs, rank, world_size)
    local = {i: segment_fn(*contigs[i]) for i in mine}
    return gather_in_order(local, costs, rank, world_size, dist, shard_dir)
