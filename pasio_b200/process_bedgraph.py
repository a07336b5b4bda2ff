"""Bedgraph in, segments out (reference: /root/reference/src/pasio/process_bedgraph.py:9-92).

Same functions and semantics as the reference: intervals are grouped per chromosome (consecutive lines),
gaps between intervals are zero-filled unless `split_at_gaps`, a contig starts at its first interval's
start, three output modes with the reference's %-formats.  What differs is where the time goes: the text
is parsed by one C++ pass (csrc/textio.cpp) instead of a per-line Python loop, contigs are handed to the
GPU as run-length intervals (pasio_contig_load_rle expands them on the device: the dense 8 B/nt profile
never exists on the host), and the output lines are formatted in C++.
"""
import numpy as np

from . import _native
from .logging import logger
from .segmentation import segments_with_scores, run_loaded_pipeline
from .dto.intervals import BedgraphInterval
from .utils.gzip_utils import open_for_read, open_for_write
from .splitters import _fusion

OUTPUT_MODES = {'bedgraph': 0, 'bed': 1, 'bedgraph+length+LMM': 2}


def fill_interval_gaps(intervals):
    """Reference helper kept for API compatibility: yields the intervals with zero-coverage gap tuples."""
    previous_stop = None
    for interval in intervals:
        start = interval[1]
        if previous_stop and previous_stop != start:
            yield (interval[0], previous_stop, start, 0)
        yield interval
        previous_stop = interval[2]


def intervals_not_adjacent(interval_1, interval_2):
    return interval_1.stop != interval_2.start


def interval_groups(intervals, split_at_gaps):
    """Reference helper kept for API compatibility (groups of BedgraphInterval objects forming one contig)."""
    import itertools
    from .utils.slice_when import slice_when
    for _, chromosome_intervals in itertools.groupby(intervals, key=lambda interval: interval.chrom):
        if split_at_gaps:
            for group in slice_when(chromosome_intervals, condition=intervals_not_adjacent):
                yield group
        else:
            yield fill_interval_gaps(chromosome_intervals)


def _read_all(stream):
    # a text stream over a binary one (open(..., 'rt'), gzip.open(..., 'rt'), sys.stdin): take the bytes underneath
    # instead of decoding gigabytes of ASCII to str and encoding them back
    raw = getattr(stream, 'buffer', None)
    if raw is not None and hasattr(raw, 'read'):
        try:
            data = raw.read()
            if isinstance(data, (bytes, bytearray)):
                return bytes(data)
        except (OSError, ValueError):
            pass
    data = stream.read()
    return data.encode() if isinstance(data, str) else bytes(data)


def contig_runs(data, split_at_gaps=False):
    """Parse bedgraph bytes and yield (chrom, run_lengths, run_values, chrom_start) per contig.

    Restatement (csrc/textio.cpp: pasio_bedgraph_runs) of interval_groups + the accumulation loop of parse_bedgraph_stream
    (reference process_bedgraph.py:26-60): consecutive lines of one chromosome form a group; with
    split_at_gaps a group also ends where an interval does not start at the previous stop; otherwise a
    zero run is inserted between non-adjacent intervals (when the previous stop is non-zero, as in the
    reference's `if previous_stop and ...`).  Runs of non-positive length contribute nothing."""
    rec = _native.parse_bedgraph_text(data)
    n = len(rec['starts'])
    if rec['n_float']:
        logger.warning("Pasio cannot be used with floating point counts. %d count(s) were automatically converted "
                       "to integers as an approximation. Make sure these values were designed to actually be "
                       "integer counts." % rec['n_float'])
    if n == 0:
        return
    run_len, run_val, group_line, group_run = _native.bedgraph_runs(rec, split_at_gaps)
    name_off, name_len = rec['name_off'][group_line].tolist(), rec['name_len'][group_line].tolist()
    first_start = rec['starts'][group_line].tolist()
    bounds = group_run.tolist()
    for k in range(len(group_line)):
        chrom = data[name_off[k]:name_off[k] + name_len[k]].decode()
        yield chrom, run_len[bounds[k]:bounds[k + 1]], run_val[bounds[k]:bounds[k + 1]], first_start[k]


def parse_bedgraph(filename, split_at_gaps=False):
    """yields (chrom, dense int profile, chromosome_start); like the reference, ignores split_at_gaps"""
    with open_for_read(filename) as stream:
        for item in parse_bedgraph_stream(stream):
            yield item


def parse_bedgraph_stream(input_stream, split_at_gaps=False):
    for chrom, run_len, run_val, chrom_start in contig_runs(_read_all(input_stream), split_at_gaps):
        yield chrom, np.repeat(run_val.astype(int), run_len), chrom_start


def split_bedgraph(in_filename, out_filename, splitter, split_at_gaps=False, output_mode='bedgraph'):
    with open_for_write(out_filename) as output_stream:
        with open_for_read(in_filename) as input_stream:
            split_bedgraph_stream(input_stream, output_stream, splitter,
                                  split_at_gaps=split_at_gaps, output_mode=output_mode)


def _write(output_stream, payload):
    raw = getattr(output_stream, 'buffer', None)
    if raw is not None and not getattr(output_stream, 'closed', False):
        output_stream.flush()
        raw.write(payload)
    else:
        try:
            output_stream.write(payload.decode('ascii'))
        except TypeError:
            output_stream.write(payload)


# Short contigs (scaffolds, transcripts) are segmented many per launch: consecutive contigs are collected until the
# batch holds BATCH_NT positions or BATCH_CONTIGS contigs and go to the device as one "super-contig" with forced
# boundaries (pasio_contig_load_rle with offsets); contigs never interact (reference process_bedgraph.py:69 handles
# them one by one), so the output is the same, in input order.  A contig of BATCH_ALONE positions or more runs alone.
BATCH_NT = 1 << 27
BATCH_CONTIGS = 20000
BATCH_ALONE = 1 << 24


def _segment_runs_on_device(plan, contigs, want_lmm):
    """contigs: list of (run_len, run_val), segmented in one launch sequence.
    -> (splits, means, lmm or None, first_split, offsets): split positions of the concatenation (they contain every
    contig boundary), per-segment outputs, the index of each contig's first split point, the contig offsets"""
    eng = _native.engine()
    eng.use_scorer(plan['factory'])
    lengths = np.array([int(rl.sum()) for rl, _ in contigs], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    if len(contigs) == 1:
        run_len, run_val = contigs[0]
        eng.load_rle(np.concatenate([[0], np.cumsum(run_len)]), run_val)
    else:
        starts = np.concatenate([[0], np.cumsum(np.concatenate([rl for rl, _ in contigs]))]).astype(np.int64)
        eng.load_rle(starts, np.concatenate([rv for _, rv in contigs]), offsets=offsets)
    _, splits, means, lmm, _ = run_loaded_pipeline(eng, plan, want_lmm=want_lmm)
    first_split = np.searchsorted(splits, offsets)           # every contig boundary is a split point
    assert np.array_equal(splits[first_split], offsets)
    return splits, means, lmm, first_split, offsets


def split_bedgraph_stream(input_stream, output_stream, splitter, split_at_gaps=False, output_mode='bedgraph'):
    logger.info('Reading input file')
    plan = _fusion.pipeline_plan(splitter)
    if output_mode not in OUTPUT_MODES:
        raise ValueError('Unknown output mode `%s`' % output_mode)
    mode = OUTPUT_MODES[output_mode]
    pending, pending_nt = [], 0          # (chrom, chrom_start, run_len, run_val) waiting for a batched launch

    def flush():
        if not pending:
            return
        splits, means, lmm, first_split, offsets = _segment_runs_on_device(
            plan, [(rl, rv) for _, _, rl, rv in pending], want_lmm=(mode == 2))
        shifts = np.array([cs for _, cs, _, _ in pending], dtype=np.int64) - offsets[:-1]
        _write(output_stream, _native.format_segments_batch([c for c, _, _, _ in pending], shifts, first_split, splits,
                                                            means if mode != 1 else None, lmm if mode == 2 else None, mode))
        for chrom, _, _, _ in pending:
            logger.info('Output of chromosome %s finished' % chrom)
        del pending[:]

    for chrom, run_len, run_val, chrom_start in contig_runs(_read_all(input_stream), split_at_gaps):
        n = int(run_len.sum())
        logger.info('Starting chrom %s of length %d' % (chrom, n))
        if plan is not None:
            # canonical splitter graph: run-length intervals go straight to the device
            assert n > 0
            if n >= BATCH_ALONE or BATCH_NT <= 0 or plan['final'] != 'nop':      # (the exact DP is a single-contig kernel)
                flush()
                pending_nt = 0
                pending.append((chrom, chrom_start, run_len, run_val))
                flush()
                continue
            if pending and (pending_nt + n > BATCH_NT or len(pending) >= BATCH_CONTIGS):
                flush()
                pending_nt = 0
            pending.append((chrom, chrom_start, run_len, run_val))
            pending_nt += n
            continue
        counts = np.repeat(run_val.astype(int), run_len)
        segs = list(segments_with_scores(counts, splitter))
        splits = np.array([s.start for s in segs] + [segs[-1].stop], dtype=np.int64)
        means = np.array([s.mean_count for s in segs], dtype=np.float64)
        lmm = np.array([s.log_marginal_likelyhood for s in segs], dtype=np.float64)
        _write(output_stream, _native.format_segments(chrom, chrom_start, splits, means if mode != 1 else None,
                                                      lmm if mode == 2 else None, mode))
        logger.info('Output of chromosome %s finished' % chrom)
    flush()
