from .square_splitter import SquareSplitter
from .nop_splitter import NopSplitter
from .constants_reducer import NotConstantReducer, NotZeroReducer
from .sliding_window_reducer import SlidingWindowReducer
from .round_reducer import RoundReducer
from .reducer_combiner import ReducerCombiner
from .default_splitters import configure_splitter
