"""Sequential composition of reducers (reference: /root/reference/src/pasio/splitters/reducer_combiner.py:1-24)."""


class ReducerCombiner(object):
    def __init__(self, *reducers):
        self.reducers = reducers

    def reduce_candidate_list(self, counts, split_candidates):
        for reducer in self.reducers:
            split_candidates = reducer.reduce_candidate_list(counts, split_candidates)
        return split_candidates

    def _final(self, method, complaint):
        last = self.reducers[-1]
        fn = getattr(last, method, None)
        if not callable(fn):
            raise Exception(complaint)
        return fn

    def split(self, counts, split_candidates):
        final_split = self._final('split', 'This ReducerCombiner has no splitter at the end of pipeline. '
                                           'Splitting no possible')
        for reducer in self.reducers[:-1]:
            split_candidates = reducer.reduce_candidate_list(counts, split_candidates)
        return final_split(counts, split_candidates)

    def scorer(self, counts, split_candidates):
        final_scorer = self._final('scorer', 'This ReducerCombiner has no splitter at the end of pipeline. '
                                             'Scoring not possible. Consider use of NopSplitter')
        return final_scorer(counts, split_candidates)
