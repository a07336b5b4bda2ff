"""ctypes front-end of oracle/dp_oracle.c -- TEST INFRASTRUCTURE ONLY.

Gives tests and bench.py's cpu_baseline a fast checker for full-size windows.  The C
code restates the same reference lines as oracle/pasio_oracle.py (cited there and in
dp_oracle.c) and is itself checked against pasio_oracle.py and the golden vectors.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import pasio_oracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libdp_oracle.so')
_lib = None

CONSTRAINTS = {'none': 0, 'zeros': 1, 'constants': 2}


def build():
    src = os.path.join(_HERE, 'dp_oracle.c')
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '_build/libdp_oracle.so'],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        _lib.dp_oracle.restype = ctypes.c_int
        _lib.dp_oracle.argtypes = [i64p, i64p, ctypes.c_int64, f64p, ctypes.c_int64, f64p,
                                   ctypes.c_int64, ctypes.c_int, ctypes.c_double,
                                   ctypes.c_double, f64p, i64p]
        _lib.round_oracle.restype = ctypes.c_int
        _lib.round_oracle.argtypes = [i64p, u8p, ctypes.c_int64, i64p, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                      f64p, ctypes.c_int64, f64p, ctypes.c_int64,
                                      ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                      u8p, i64p]
        _lib.dp_oracle_mt.restype = ctypes.c_int
        _lib.dp_oracle_mt.argtypes = _lib.dp_oracle.argtypes + [ctypes.c_int]
        _lib.round_oracle_mt.restype = ctypes.c_int
        _lib.round_oracle_mt.argtypes = _lib.round_oracle.argtypes + [ctypes.c_int]
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


class FlatOracle(object):
    """Whole-contig state for the flat formulation: global cumsum,
    change-point flags and host tables extended on demand."""

    def __init__(self, counts, alpha, beta, threads=1):
        """threads > 1: columns of a DP row / windows of a round spread over OpenMP threads (same cells, same
        first-maximum rule; test_oracle_golden.py checks mt == single-threaded)."""
        self.threads = int(threads)
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        assert len(counts) > 0 and np.all(counts >= 0)
        self.counts = counts
        self.n = len(counts)
        self.alpha = po.normalise_alpha(alpha)
        self.beta = beta
        self.int_alpha = isinstance(self.alpha, (int, np.integer))
        self.tables = po.Tables(self.alpha, beta, cache_size=2)
        self.Cg = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.cp = np.zeros(self.n + 1, dtype=np.uint8)
        self.cp[1:self.n] = counts[1:] != counts[:-1]
        self.lg = self.g = None
        # log_marginal_likelyhood.py:62
        self.pen = float(self.alpha * np.log(0 + beta) - __import__('scipy.special').special.gammaln(0 + self.alpha))

    def _ensure_tables(self, n_log, n_gam):
        if self.lg is None or len(self.lg) < n_log or len(self.g) < n_gam:
            n_log = max(n_log, 0 if self.lg is None else len(self.lg))
            n_gam = max(n_gam, 0 if self.g is None else len(self.g))
            lg, g, ga = po.extended_tables(self.tables, n_log, n_gam)
            self.lg = np.ascontiguousarray(lg)
            self.g = np.ascontiguousarray(g if self.int_alpha else ga)

    def square_split(self, cands):
        """Exact DP; returns (score, splits, P, prev)."""
        cands = np.ascontiguousarray(cands, dtype=np.int64)
        C = np.ascontiguousarray(self.Cg[cands])
        N = len(cands)
        a_int = int(self.alpha) if self.int_alpha else 0
        self._ensure_tables(int(cands[-1] - cands[0]) + 1, int(C[-1] - C[0]) + a_int + 1)
        P = np.empty(N)
        prev = np.empty(N, dtype=np.int64)
        args = (_p(C, ctypes.c_int64), _p(cands, ctypes.c_int64), N,
                _p(self.g, ctypes.c_double), len(self.g),
                _p(self.lg, ctypes.c_double), len(self.lg),
                int(self.int_alpha), float(self.alpha), self.pen,
                _p(P, ctypes.c_double), _p(prev, ctypes.c_int64))
        rc = lib().dp_oracle_mt(*(args + (self.threads,))) if self.threads > 1 else lib().dp_oracle(*args)
        assert rc == 0, rc
        idx = po.backtrace(prev)
        return P[-1], cands[idx], P, prev

    def round(self, cands, window_size, window_shift, constraint):
        """One sliding-window round; returns (new candidates, cells)."""
        cands = np.ascontiguousarray(cands, dtype=np.int64)
        m = len(cands)
        span = cnt = 0
        for st, en in po.window_ranges(m, window_size, window_shift):
            a, b = cands[st], cands[en - 1]
            span = max(span, int(b - a))
            cnt = max(cnt, int(self.Cg[b] - self.Cg[a]))
        a_int = int(self.alpha) if self.int_alpha else 0
        self._ensure_tables(span + 1, cnt + a_int + 1)
        keep = np.zeros(self.n + 1, dtype=np.uint8)
        cells = ctypes.c_int64(0)
        args = (_p(self.Cg, ctypes.c_int64), _p(self.cp, ctypes.c_uint8), self.n,
                _p(cands, ctypes.c_int64), m, window_size, window_shift,
                CONSTRAINTS[constraint],
                _p(self.g, ctypes.c_double), len(self.g),
                _p(self.lg, ctypes.c_double), len(self.lg),
                int(self.int_alpha), float(self.alpha), self.pen,
                _p(keep, ctypes.c_uint8), ctypes.byref(cells))
        rc = lib().round_oracle_mt(*(args + (self.threads,))) if self.threads > 1 else lib().round_oracle(*args)
        assert rc == 0, rc
        return np.flatnonzero(keep).astype(np.int64), cells.value

    def rounds(self, window_size, window_shift, constraint, num_rounds=None, cands=None):
        """round_reducer.py:10-31 over the flat round; returns (cands, sizes, cells)."""
        if cands is None:
            cands = np.arange(self.n + 1, dtype=np.int64)
        limit = max(1, self.n if num_rounds is None else num_rounds)
        sizes, total_cells = [len(cands)], 0
        for _ in range(limit):
            new, cells = self.round(cands, window_size, window_shift, constraint)
            total_cells += cells
            if len(new) == len(cands):
                return new, sizes, total_cells
            cands = new
            sizes.append(len(cands))
        return cands, sizes, total_cells
