"""Host-side builders of the log / log-gamma look-up tables.

Interface of the reference's /root/reference/src/pasio/cached_log.py:5-56 (`LogComputer`,
`LogGammaComputer`: same constructor, `compute_for_number`, `compute_for_array`,
`compute_for_array_unbound`).  Role here: the tables are the ONLY place transcendental values
come from -- they are produced by the same `np.log` / `scipy.special.gammaln` calls the
reference makes and uploaded to the GPU, where the DP kernels gather from them.  Arguments
past `cache_size` (which the reference evaluates on the fly, cached_log.py:24-29,51-56) are
served by extending the table with the same calls; SURVEY 7.3 checked that an extended table
is bit-identical to the reference's direct evaluation.
"""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np
import scipy.special

_CHUNK = 1 << 20


def _build(fn, shift, lo, hi):
    """fn(arange(lo, hi) + shift) evaluated in chunks on a few threads (ufunc loops drop the GIL)."""
    if hi - lo <= 2 * _CHUNK:
        return fn(np.arange(lo, hi) + shift)
    edges = list(range(lo, hi, _CHUNK)) + [hi]
    workers = min(16, os.cpu_count() or 1)
    with ThreadPoolExecutor(workers) as pool:
        parts = list(pool.map(lambda ab: fn(np.arange(ab[0], ab[1]) + shift), zip(edges[:-1], edges[1:])))
    return np.concatenate(parts)


class _TableComputer(object):
    _fn = None

    def __init__(self, shift=0, cache_size=1048576):
        self.cache_size = cache_size
        self.shift = shift
        self.precomputed = type(self)._fn(np.arange(self.cache_size) + shift)
        self._extended = self.precomputed

    def __reduce__(self):
        # pickled as its constructor arguments (device_pool.py sends splitters to worker processes): the tables are
        # rebuilt by the same numpy / scipy calls on the other side, not shipped
        return (type(self), (self.shift, self.cache_size))

    def table(self, n):
        """float64 table of at least n entries; entry k is fn(k + shift)."""
        if n > len(self._extended):
            with np.errstate(all='ignore'):
                tail = _build(type(self)._fn, self.shift, len(self._extended), int(n))
            self._extended = np.ascontiguousarray(np.concatenate([self._extended, tail]))
        return self._extended

    def compute_for_number(self, x):
        if x < self.cache_size:
            return self.precomputed[x]
        return type(self)._fn(x + self.shift)

    def compute_for_array(self, x, max_value):
        if max_value < self.cache_size:
            return self.precomputed[x]
        return self.compute_for_array_unbound(x)

    def compute_for_array_unbound(self, x):
        x = np.asarray(x)
        top = int(x.max()) + 1 if x.size else 0
        return self.table(max(top, self.cache_size))[x]


class LogComputer(_TableComputer):
    """log(k + shift), k = 0..cache_size-1 (reference cached_log.py:5-29)."""
    _fn = staticmethod(np.log)

    def __init__(self, shift=0, cache_size=1048576):
        with np.errstate(divide='ignore'):
            _TableComputer.__init__(self, shift, cache_size)


class LogGammaComputer(_TableComputer):
    """gammaln(k + shift), k = 0..cache_size-1 (reference cached_log.py:32-56)."""
    _fn = staticmethod(scipy.special.gammaln)
