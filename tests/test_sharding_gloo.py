"""N>1 path on CPU: two gloo ranks partition contigs by LPT and rank 0 gathers the per-contig results
in input order (pasio_b200/sharding.py).  The per-contig work is stubbed with the oracle so the test
needs no GPU; on the GPU box the same plumbing carries segment_on_device results."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pasio_b200 import sharding, synth
    from oracle import c_oracle
    lens = [3000, 12000, 500, 7000, 2500, 9000, 100]
    contigs = [('ctg%d' % k, synth.dnase_like(n, 50 + k, hotspot_share=0.5), 10 * k) for k, n in enumerate(lens)]

    def segment(name, counts, start):
        splits, _, _ = c_oracle.FlatOracle(counts, 1.0, 1.0).rounds(200, 100, 'constants')
        return ''.join('%s\t%d\t%d\n' % (name, a + start, b + start) for a, b in zip(splits[:-1], splits[1:]))

    mine = sharding.shard_indices([len(c[1]) for c in contigs], rank, world)
    res = sharding.segment_contigs(contigs, segment, rank=rank, world_size=world, dist=dist)
    shard_dir = os.path.join(out_dir, 'shards')
    os.makedirs(shard_dir, exist_ok=True)
    res_files = sharding.segment_contigs(contigs, segment, rank=rank, world_size=world, dist=dist, shard_dir=shard_dir)
    if rank == 0:
        serial = [segment(*c) for c in contigs]
        assert res == serial
        assert res_files == serial            # the same through per-rank shard files
        with open(os.path.join(out_dir, 'ok'), 'w') as f:
            f.write(','.join(map(str, mine)))
    else:
        assert res is None and res_files is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_lpt_and_gather(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    mine0 = (tmp_path / 'ok').read_text()
    assert mine0 != ''          # rank 0 had work and the gathered text equalled the serial run
