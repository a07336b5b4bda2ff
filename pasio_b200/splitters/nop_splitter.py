"""NopSplitter: keeps the candidates, scores them (reference: /root/reference/src/pasio/splitters/nop_splitter.py:5-18)."""
import numpy as np


class NopSplitter(object):
    def __init__(self, scorer_factory):
        self.scorer_factory = scorer_factory

    def scorer(self, counts, split_candidates):
        return self.scorer_factory(counts, split_candidates)

    def reduce_candidate_list(self, counts, split_candidates):
        return split_candidates

    def split(self, counts, split_candidates):
        # per-segment scores come from the device (csrc/score.cu); np.sum is the reference's own
        # pairwise reduction of those m-1 numbers, kept so the total rounds identically
        scores = self.scorer(counts, split_candidates).scores()
        return (np.sum(scores), split_candidates)
