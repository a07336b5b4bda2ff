"""Scorer objects: the reference's log-marginal-likelihood interface, computed on the GPU.

Mirrors /root/reference/src/pasio/log_marginal_likelyhood.py (ScorerFactory :5-28,
LogMarginalLikelyhoodComputer :45-99, IntAlpha :101-115, RealAlpha :117-132): same class
names, constructor arguments, methods and attributes, same AssertionErrors.  The arrays
(`cumsum`, row scores, per-segment scores, means, log-factorial sums) are produced by the
sm_100a kernels behind include/pasio_b200.h; only scalar bookkeeping happens here.
"""
from __future__ import division

import numpy as np

from . import _native
from .cached_log import LogComputer, LogGammaComputer


def _normalise_alpha(alpha):
    # reference :9-12 -- integral floats select the integer-alpha scorer
    if isinstance(alpha, float) and alpha.is_integer():
        return int(alpha)
    return alpha


def _creation_cost(alpha, log_computer, log_gamma_alpha_computer):
    # reference :62 -- alpha*log(beta) - gammaln(alpha), both read from entry 0 of the tables
    with np.errstate(all='ignore'):
        return alpha * log_computer.compute_for_number(0) - log_gamma_alpha_computer.compute_for_number(0)


class ScorerFactory(object):
    def __init__(self, alpha, beta):
        assert alpha >= 0
        assert beta >= 0
        self.alpha = _normalise_alpha(alpha)
        self.beta = beta
        self.log_gamma_computer = LogGammaComputer()
        self.log_gamma_alpha_computer = LogGammaComputer(shift=alpha)
        self.log_computer = LogComputer(shift=beta)
        self.segment_creation_cost = _creation_cost(self.alpha, self.log_computer, self.log_gamma_alpha_computer)

    def __reduce__(self):
        return (ScorerFactory, (self.alpha, self.beta))      # tables are rebuilt, not pickled

    def __call__(self, counts, split_candidates):
        cls = (LogMarginalLikelyhoodIntAlphaComputer if isinstance(self.alpha, int)
               else LogMarginalLikelyhoodRealAlphaComputer)
        return cls(counts, self.alpha, self.beta, split_candidates,
                   log_computer=self.log_computer,
                   log_gamma_computer=self.log_gamma_computer,
                   log_gamma_alpha_computer=self.log_gamma_alpha_computer)


def assert_correct_counts(counts):
    assert isinstance(counts, np.ndarray)
    assert counts.dtype == int
    assert len(counts) > 0
    # counts >= 0 is checked by the scan kernel while it reads the data (PASIO_E_COUNTS)


def assert_correct_split_candidates(split_candidates, counts):
    assert isinstance(split_candidates, np.ndarray)
    assert len(split_candidates) >= 1
    assert split_candidates[0] == 0
    assert split_candidates[-1] == len(counts)
    # strictly ascending is checked on the device (PASIO_E_CANDIDATES)


def _is_all_positions(split_candidates, n):
    if len(split_candidates) != n + 1:
        return False
    return bool(np.all(split_candidates[1:] - split_candidates[:-1] == 1))


class LogMarginalLikelyhoodComputer(object):
    """Indexing runs over split candidates, not counts (as in the reference)."""

    def __init__(self, counts, alpha, beta, split_candidates,
                 log_computer=None, log_gamma_computer=None, log_gamma_alpha_computer=None):
        self.alpha = alpha
        self.log_computer = log_computer if log_computer else LogComputer(shift=beta)
        self.log_gamma_computer = log_gamma_computer if log_gamma_computer else LogGammaComputer()
        self.log_gamma_alpha_computer = (log_gamma_alpha_computer if log_gamma_alpha_computer
                                         else LogGammaComputer(shift=alpha))
        assert_correct_counts(counts)
        assert_correct_split_candidates(split_candidates, counts)
        self.split_candidates = split_candidates
        self._counts = counts
        self._implicit = _is_all_positions(split_candidates, len(counts))
        self.segment_creation_cost = _creation_cost(alpha, self.log_computer, self.log_gamma_alpha_computer)
        self._cumsum = None
        self._logfac_cumsum = None
        self._bind()      # loads + validates on the device now, like the reference's constructor asserts

    # the engine holds one contig and one candidate list at a time; re-binding is cached by identity
    def _bind(self):
        eng = _native.engine()
        eng.use_scorer(self)
        eng.load(self._counts)
        eng.set_candidates(None if self._implicit else self.split_candidates)
        return eng

    @property
    def cumsum(self):
        if self._cumsum is None:
            self._cumsum = self._bind().cumsum_at_candidates()
        return self._cumsum

    @property
    def logfac_cumsum(self):
        if self._logfac_cumsum is None:
            self._logfac_cumsum = self._bind().segment_scores(scores=False, logfac=True)[3]
        return self._logfac_cumsum

    def total_sum_logfac(self):
        return self.logfac_cumsum[-1]

    def scores(self):
        return self._bind().segment_scores(scores=True)[0]

    def log_marginal_likelyhoods(self):
        return self.scores() - np.diff(self.logfac_cumsum)

    def mean_counts(self):
        return self._bind().segment_scores(scores=False, means=True)[2]

    def score(self, start, stop):
        return self.self_score(start, stop) + self.segment_creation_cost

    def self_score(self, start, stop):
        # scalar bookkeeping on two table entries (reference :88-94)
        segment_count = self.cumsum[stop] - self.cumsum[start]
        shifted_segment_count = segment_count + self.alpha
        segment_length = self.split_candidates[stop] - self.split_candidates[start]
        add = self.log_gamma_alpha_computer.compute_for_number(segment_count)
        sub = shifted_segment_count * self.log_computer.compute_for_number(segment_length)
        return add - sub

    def self_score_no_splits(self):
        return self.self_score(0, len(self.split_candidates) - 1)

    def score_no_splits(self):
        return self.self_score_no_splits() + self.segment_creation_cost

    def all_suffixes_self_score(self, stop):
        """Scores of segments [i, stop) for all i < stop (reference :105-115 / :121-132)."""
        return self._bind().suffix_scores(int(stop))

    # exact DP over this scorer's candidates (SquareSplitter.split_without_normalizations)
    def _square_split(self):
        score, splits = self._bind().square_split()
        return score, splits

    # regularised DP over this scorer's candidates (SquareSplitter.split_with_normalizations) with host-built penalty tables
    def _square_split_regularized(self, length_penalty, split_number_penalty, first_column_refund):
        return self._bind().square_split_regularized(length_penalty, split_number_penalty, first_column_refund)


class LogMarginalLikelyhoodIntAlphaComputer(LogMarginalLikelyhoodComputer):
    pass


class LogMarginalLikelyhoodRealAlphaComputer(LogMarginalLikelyhoodComputer):
    pass
