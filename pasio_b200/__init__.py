"""pasio_b200 -- B200-native (sm_100a) drop-in for the segmentation hot path of autosome-ru/pasio.

Same public surface as the reference package (/root/reference/src/pasio/__init__.py:1-4).
"""
from .splitters import configure_splitter
from .segmentation import segments_with_scores
from .process_bedgraph import parse_bedgraph, split_bedgraph
from .version import __version__


def reset_device_cache():
    """Forget which counts / candidate arrays are resident on the GPU.  The engine recognises arrays it has
    already uploaded by object identity; call this after mutating such an array in place."""
    from . import _native
    if _native._engine is not None:
        _native._engine.invalidate()
