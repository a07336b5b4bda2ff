"""configure_splitter: build the splitter object graph from CLI-style options.

Signature and validation of /root/reference/src/pasio/splitters/default_splitters.py:12-66.
Two reference crashes are fixed consciously: `length_regularization_function='revlog'`
(NameError: numpy never imported there) and `algorithm='slidingwindow'` (NameError: `splitter`),
which now builds the graph the reference's own test uses for that mode
(tests/test_pasio.py:263-270): ReducerCombiner(SlidingWindowReducer, SquareSplitter).
"""
import numpy as np

from .square_splitter import SquareSplitter, _identity, _revlog
from .nop_splitter import NopSplitter
from .constants_reducer import NotConstantReducer, NotZeroReducer
from .sliding_window_reducer import SlidingWindowReducer
from .round_reducer import RoundReducer
from .reducer_combiner import ReducerCombiner
from ..dto.sliding_window import SlidingWindow
from ..log_marginal_likelyhood import ScorerFactory

# module-level functions (not lambdas): splitters must pickle for the per-GPU worker processes, and the device
# version of the regularised DP recognises exactly these two
REGULARIZATION_FUNCTIONS = {
    'none': _identity,
    'revlog': _revlog,
}


# all unknown arguments (kwargs) are ignored, like in the reference
def configure_splitter(alpha=1, beta=1, algorithm='rounds',
                       window_size=2500, window_shift=1250, num_rounds=None,
                       split_constraints='constants',
                       length_regularization_function='none', length_regularization=0, split_number_regularization=0,
                       **kwargs):
    if algorithm not in ('exact', 'slidingwindow', 'rounds'):
        raise ValueError('Algorithm should be one of exact/slidingwindow/rounds')
    windowed = algorithm in ('slidingwindow', 'rounds')
    if windowed and window_shift is None:
        raise ValueError('Argument window_shift is required for algorithms slidingwingow and rounds')
    if windowed and window_size is None:
        raise ValueError('Argument window_size is required for algorithms slidingwingow and rounds')
    if length_regularization != 0 and length_regularization_function == 'none':
        raise ValueError('Argument --length_regularization_function is required '
                         'for length regularization multiplier %s' % length_regularization)
    if length_regularization_function != 'none' and length_regularization == 0:
        raise ValueError('Argument --length_regularization_multiplier is required '
                         'for length legularization function %s' % length_regularization_function)
    if windowed and split_constraints not in ('constants', 'zeros', 'none'):
        raise ValueError('Unknown split_constraints option `%s`' % split_constraints)

    scorer_factory = ScorerFactory(alpha, beta)
    square_splitter = SquareSplitter(
        scorer_factory,
        length_regularization_multiplier=length_regularization,
        length_regularization_function=REGULARIZATION_FUNCTIONS[length_regularization_function],
        split_number_regularization_multiplier=split_number_regularization,
        split_number_regularization_function=REGULARIZATION_FUNCTIONS['none'])
    if algorithm == 'exact':
        return square_splitter

    if split_constraints == 'constants':
        base_splitter = ReducerCombiner(NotConstantReducer(), square_splitter)
    elif split_constraints == 'zeros':
        base_splitter = ReducerCombiner(NotZeroReducer(), square_splitter)
    else:
        base_splitter = square_splitter
    window_reducer = SlidingWindowReducer(sliding_window=SlidingWindow(window_size=window_size, window_shift=window_shift),
                                          base_reducer=base_splitter)
    if algorithm == 'slidingwindow':
        return ReducerCombiner(window_reducer, square_splitter)
    reducer = RoundReducer(base_reducer=window_reducer, num_rounds=num_rounds)
    return ReducerCombiner(reducer, NopSplitter(scorer_factory))
