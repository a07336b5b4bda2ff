"""pasio_b200 -- B200-native (sm_100a) drop-in for the segmentation hot path of autosome-ru/pasio.

Same public surface as the reference package (/root/reference/src/pasio/__init__.py:1-4).
"""
from .splitters import configure_splitter
from .segmentation import segments_with_scores
from .process_bedgraph import parse_bedgraph, split_bedgraph
from .version import __version__
