"""NotZeroReducer / NotConstantReducer.

Interface of /root/reference/src/pasio/splitters/constants_reducer.py:4-21.  Inside the fused
sliding-window rounds these rules run per window on the GPU (stream compaction in csrc/window_dp.cu over
the change-point bitmap of csrc/scan.cu).  Called on their own with a coverage profile (int ndarray)
they run on the GPU as well (pasio_filter_candidates: bitmap test + ordered compaction).  Anything
else -- the reference's tests also pass plain sequences -- is not a coverage profile the device can hold
and is handled with the reference's own array expressions.
"""
import numpy as np

from .. import _native
from ..logging import logger


def _is_profile(counts, split_candidates):
    return (isinstance(counts, np.ndarray) and counts.dtype == int and counts.ndim == 1 and len(counts) > 0
            and isinstance(split_candidates, np.ndarray) and len(split_candidates) >= 2
            and split_candidates[0] == 0 and split_candidates[-1] == len(counts))


def _device_filter(counts, split_candidates, constraint):
    from .sliding_window_reducer import _set_candidates
    eng = _native.engine()
    eng.load(counts)
    _set_candidates(eng, counts, split_candidates)
    n_in, n_out = eng.filter_candidates(constraint)
    return eng.candidates(), n_in, n_out


class NotZeroReducer(object):
    constraint = 'zeros'

    def reduce_candidate_list(self, counts, split_candidates):
        if _is_profile(counts, split_candidates):
            reduced, n_in, n_out = _device_filter(counts, split_candidates, 'zeros')
            if n_out != n_in:
                logger.info('Just zeros: %d --> 2 split points' % n_in)
                return reduced
            logger.info('Not zeros. Interval not reduced.')
            return split_candidates
        if not np.any(counts):
            logger.info('Just zeros: %d --> 2 split points' % len(split_candidates))
            return np.array([0, len(counts)])
        logger.info('Not zeros. Interval not reduced.')
        return split_candidates


class NotConstantReducer(object):
    constraint = 'constants'

    def reduce_candidate_list(self, counts, split_candidates):
        if _is_profile(counts, split_candidates):
            reduced, n_in, n_out = _device_filter(counts, split_candidates, 'constants')
        else:
            counts = np.asarray(counts)
            cands = np.asarray(split_candidates)
            inner = cands[(cands > 0) & (cands < len(counts))]
            keep = inner[counts[inner - 1] != counts[inner]]
            reduced = np.concatenate([[0], keep, [len(counts)]]).astype(int)
        logger.info('Constants reduced: %d --> %d split points' % (len(split_candidates), len(reduced)))
        return reduced
