"""NotZeroReducer / NotConstantReducer as stand-alone reducers.

Interface of /root/reference/src/pasio/splitters/constants_reducer.py:4-21.  Inside the fused
sliding-window rounds these rules run on the GPU (warp-ballot compaction in csrc/window_dp.cu over
the change-point bitmap of csrc/scan.cu); the methods below serve direct calls on host arrays.
"""
import numpy as np

from ..logging import logger


class NotZeroReducer(object):
    constraint = 'zeros'

    def reduce_candidate_list(self, counts, split_candidates):
        if not np.any(counts):
            logger.info('Just zeros: %d --> 2 split points' % len(split_candidates))
            return np.array([0, len(counts)])
        logger.info('Not zeros. Interval not reduced.')
        return split_candidates


class NotConstantReducer(object):
    constraint = 'constants'

    def reduce_candidate_list(self, counts, split_candidates):
        counts = np.asarray(counts)
        # a candidate p survives when counts[p-1] != counts[p]; both ends always survive
        cands = np.asarray(split_candidates)
        inner = cands[(cands > 0) & (cands < len(counts))]
        keep = inner[counts[inner - 1] != counts[inner]]
        reduced = np.concatenate([[0], keep, [len(counts)]]).astype(cands.dtype if cands.size else int)
        logger.info('Constants reduced: %d --> %d split points' % (len(split_candidates), len(reduced)))
        return reduced
