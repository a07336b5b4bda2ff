"""Bedgraph in, segments out (reference: /root/reference/src/pasio/process_bedgraph.py:9-92).

Same functions and semantics: gaps between intervals are zero-filled unless `split_at_gaps`,
a contig starts at its first interval's start, three output modes with the reference's %-formats.
The dense profile is built with np.repeat instead of a Python list.
"""
import itertools

import numpy as np

from .logging import logger
from .utils.slice_when import slice_when
from .segmentation import segments_with_scores
from .dto.intervals import BedgraphInterval
from .utils.gzip_utils import open_for_read, open_for_write


def fill_interval_gaps(intervals):
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
    """Groups of intervals that form one contig each: per chromosome, and additionally cut at
    uncovered positions when `split_at_gaps`; otherwise inner gaps are filled with zeros.
    Chromosome flanks are never filled (the chromosome length is unknown)."""
    for _, chromosome_intervals in itertools.groupby(intervals, key=lambda interval: interval.chrom):
        if split_at_gaps:
            for group in slice_when(chromosome_intervals, condition=intervals_not_adjacent):
                yield group
        else:
            yield fill_interval_gaps(chromosome_intervals)


def parse_bedgraph(filename, split_at_gaps=False):
    """yields (chrom, dense int profile, chromosome_start); like the reference, ignores split_at_gaps"""
    with open_for_read(filename) as stream:
        for item in parse_bedgraph_stream(stream):
            yield item


def parse_bedgraph_stream(input_stream, split_at_gaps=False):
    intervals_stream = BedgraphInterval.each_in_stream(input_stream)
    for group in interval_groups(intervals_stream, split_at_gaps=split_at_gaps):
        chromosome = chromosome_start = None
        lengths, values = [], []
        for (chrom, start, stop, coverage) in group:
            if chromosome_start is None:
                chromosome_start, chromosome = start, chrom
            lengths.append(max(stop - start, 0))
            values.append(coverage)
        profile = np.repeat(np.array(values, dtype=int), np.array(lengths, dtype=int))
        yield chromosome, profile, chromosome_start


def split_bedgraph(in_filename, out_filename, splitter, split_at_gaps=False, output_mode='bedgraph'):
    with open_for_write(out_filename) as output_stream:
        with open_for_read(in_filename) as input_stream:
            split_bedgraph_stream(input_stream, output_stream, splitter,
                                  split_at_gaps=split_at_gaps, output_mode=output_mode)


_FORMATS = {
    'bedgraph': lambda chrom, off, s: '%s\t%d\t%d\t%f\n' % (chrom, s.start + off, s.stop + off, s.mean_count),
    'bed': lambda chrom, off, s: '%s\t%d\t%d\n' % (chrom, s.start + off, s.stop + off),
    'bedgraph+length+LMM': lambda chrom, off, s: '%s\t%d\t%d\t%f\t%d\t%f\n' % (
        chrom, s.start + off, s.stop + off, s.mean_count, s.length, s.log_marginal_likelyhood),
}


def split_bedgraph_stream(input_stream, output_stream, splitter, split_at_gaps=False, output_mode='bedgraph'):
    logger.info('Reading input file')
    for chrom, counts, chrom_start in parse_bedgraph_stream(input_stream, split_at_gaps=split_at_gaps):
        logger.info('Starting chrom %s of length %d' % (chrom, len(counts)))
        if output_mode not in _FORMATS:
            raise ValueError('Unknown output mode `%s`' % output_mode)
        fmt = _FORMATS[output_mode]
        for scored_interval in segments_with_scores(counts, splitter):
            output_stream.write(fmt(chrom, chrom_start, scored_interval))
        logger.info('Output of chromosome %s finished' % chrom)
