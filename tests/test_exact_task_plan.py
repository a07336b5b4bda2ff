"""The worker task list of the exact DP (csrc/exact_pruned.cu, build_tasks) keeps the ordering that makes the kernel's flag
waits deadlock-free: worker CTAs pull tasks in list order and all CTAs are co-resident, so it suffices that every task only
waits for tasks EARLIER in the list, or for a progress of the diagonal CTA that itself needs nothing later in the list.
Host-only (pasio_exact_task_plan launches nothing): runs without a GPU."""
import ctypes

import numpy as np
import pytest

from pasio_b200 import _native

RB, G, NG, SQ, SAHEAD, RING = 128, 8, 8, 4, 16, 64
S, F, N, R = 0, 1, 2, 3


def plan(n, lag, nblock):
    lib = _native.load_library()
    cnt = ctypes.c_int64()
    assert lib.pasio_exact_task_plan(n, lag, nblock, None, 0, ctypes.byref(cnt)) != 0      # size query
    out = np.zeros(3 * max(cnt.value, 1), dtype=np.int32)
    rc = lib.pasio_exact_task_plan(n, lag, nblock, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), cnt.value, ctypes.byref(cnt))
    assert rc == 0
    return [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(cnt.value)]


def has_n(b, lag, nb):               # xp_has_n
    return nb > 0 and 1 + RB * (b - lag + 1 + nb) > 1


SHAPES = [2, 3, 100, 128, 129, 130, 257, 385, 513, 641, 769, 1000, 4101, 8321, 8322, 9000, 20001, 100001]


@pytest.mark.parametrize('lag', [3, 4, 5])
def test_task_list_order(lag):
    for nb in range(0, lag - 1):
        for n in SHAPES:
            tasks = plan(n, lag, nb)
            nB = (n - 1 + RB - 1) // RB
            where = (n, lag, nb)
            assert len(set(tasks)) == len(tasks), ('duplicate task',) + where
            assert all(0 <= t[0] < nB for t in tasks)
            slices = {}                                   # (kind, block) -> slices; first / last position
            first, last = {}, {}
            for i, (blk, kind, sl) in enumerate(tasks):
                slices.setdefault((kind, blk), []).append(sl)
                first.setdefault((kind, blk), i)
                last[(kind, blk)] = i
            # every row block has its self scores, and exactly the result slices the diagonal waits for (xp_far_count)
            for blk in range(nB):
                assert sorted(slices.get((S, blk), [])) == list(range(SQ)), where + (blk,)
                assert sorted(slices.get((F, blk), [])) == (list(range(G)) if blk >= lag - 1 else []), where + (blk,)
                assert sorted(slices.get((N, blk), [])) == (list(range(NG)) if has_n(blk, lag, nb) else []), where + (blk,)
            assert sorted(blk for (kind, blk) in slices if kind == R) == list(range(max(0, nB - lag))), where
            # latest list position of anything the diagonal needs from worker CTAs to FINISH the blocks < c: the F / N results of
            # those blocks, the self scores of the blocks <= c (the sweeping warps prefetch one block ahead)
            needs = [-1] * (nB + 2)
            for c in range(1, nB + 2):
                m = needs[c - 1]
                for kind in (F, N):
                    m = max(m, last.get((kind, c - 1), -1))
                m = max(m, last.get((S, c - 1), -1), last.get((S, c), -1))
                needs[c] = m
            for i, (blk, kind, sl) in enumerate(tasks):
                if kind == F:                            # waits for done_block >= blk - lag + 1: R(blk - lag)
                    if blk - lag >= 0:
                        assert last[(R, blk - lag)] < i, where + ('F', blk)
                elif kind == N:                          # waits for p_block >= blk - lag + 1 + nb (block by block)
                    assert needs[max(0, min(blk - lag + 1 + nb, nB))] < i, where + ('N', blk)
                elif kind == R:                          # waits for p_block >= blk + 1 and for R(blk - 1)
                    assert needs[blk + 1] < i, where + ('R', blk)
                    if blk >= 1:
                        assert last[(R, blk - 1)] < i, where + ('R order', blk)
                elif kind == S and nB > RING and blk >= RING:    # ring slot reuse: done_block >= blk - RING + 1
                    assert last[(R, blk - RING)] < i, where + ('S', blk)


def test_task_plan_rejects_bad_arguments():
    lib = _native.load_library()
    cnt = ctypes.c_int64()
    assert lib.pasio_exact_task_plan(1, 3, 1, None, 0, ctypes.byref(cnt)) != 0
    assert lib.pasio_exact_task_plan(1000, 2, 1, None, 0, ctypes.byref(cnt)) != 0
    assert lib.pasio_exact_task_plan(1000, 6, 1, None, 0, ctypes.byref(cnt)) != 0
