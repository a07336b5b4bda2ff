"""Open plain / gzipped files or stdin/stdout (reference: /root/reference/src/pasio/utils/gzip_utils.py:15-39)."""
import contextlib
import gzip
import sys


def choose_open_function(filename, force_gzip=None):
    if force_gzip not in (True, False, None):
        raise ValueError("`force_gzip` should be one of True/False/None")
    use_gzip = filename.endswith('.gz') if force_gzip is None else force_gzip
    return gzip.open if use_gzip else open


def _open(filename, force_gzip, mode, std_stream):
    if filename and filename != '-':
        return choose_open_function(filename, force_gzip)(filename, mode)
    return contextlib.nullcontext(std_stream)


def open_for_write(filename, force_gzip=None, mode='wt'):
    return _open(filename, force_gzip, mode, sys.stdout)


def open_for_read(filename, force_gzip=None, mode='rt'):
    return _open(filename, force_gzip, mode, sys.stdin)
