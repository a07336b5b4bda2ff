"""Command line front-end with the reference's flags (/root/reference/src/pasio/cli.py:9-92)."""
import argparse
import logging
import sys

from .logging import logger
from .process_bedgraph import split_bedgraph
from .splitters.default_splitters import configure_splitter
from .version import __version__


def get_argparser():
    p = argparse.ArgumentParser(
        prog="pasio",
        description='PASIO produces segmentation of coverage profile into regions with uniform coverage\n',
        usage='pasio input.bedgraph[.gz] [options]',
        formatter_class=argparse.RawTextHelpFormatter)
    p.add_argument('bedgraph', metavar='input_bedgraph', help="Input file in bedgraph format\n(it can be gzipped)")
    p.add_argument('--alpha', '-a', type=float, default=1.0, metavar='VAL',
                   help="alpha parameter of gamma distribution (default: %(default)s)")
    p.add_argument('--beta', '-b', type=float, default=1.0, metavar='VAL',
                   help="beta parameter of gamma distribution (default: %(default)s)")
    p.add_argument('--output-file', '-o', metavar='FILE', dest='output_file',
                   help="Output file. It will be in bedgraph/bed/tsv format\n(can be gzipped)")
    p.add_argument('--output-mode', metavar='MODE', default='bedgraph',
                   choices=['bedgraph', 'bedgraph+length+LMM', 'bed'],
                   help="Formatting of output. Default: %(default)s.\nPossible options: %(choices)s")
    p.add_argument('--algorithm', choices=['slidingwindow', 'exact', 'rounds'], default='rounds', metavar='ALGO',
                   help="Algorithm to use (default: %(default)s)\nPossible options: %(choices)s")
    p.add_argument('--split-constraints', metavar='STRATEGY', choices=('none', 'zeros', 'constants'),
                   default='constants',
                   help="Specify types of intervals which shouldn't be splitted.\n"
                        "Default: %(default)s\nOptions: %(choices)s")
    p.add_argument('--split-number-regularization', type=float, default=0, metavar='VALUE',
                   help="Penalty multiplier for each split")
    p.add_argument('--length-regularization', type=float, default=0, metavar='VALUE',
                   help="Penalty multiplier for length of each segment")
    p.add_argument('--length-regularization-function', type=str, default='none', metavar='FUNC',
                   choices=['none', 'revlog'],
                   help='Penalty function for length of segments:\nDefault: %(default)s. Possible options:\n'
                        '* none -- no length regulatization\n* revlog -- 1/log(1+l)\n')
    p.add_argument('--window-size', type=int, default=2500, metavar='SIZE',
                   help="Size of window for slidingwindow/rounds algorithms\n(default: %(default)s)")
    p.add_argument('--window-shift', type=int, default=1250, metavar='SHIFT',
                   help="Shift in one step (default: %(default)s)")
    p.add_argument('--num-rounds', type=int, metavar='N',
                   help='Number of rounds for round algorithm.\nIf not set, run until no split points removed')
    p.add_argument('--split-at-gaps', action='store_true',
                   help='By default gaps between intervals are filled with zeros.\n'
                        'Split at gaps overrides this behavior so that\n'
                        'non-adjacent intervals are segmented independently.')
    p.add_argument('--devices', type=int, default=None, metavar='N',
                   help='(pasio_b200) shard the contigs over N GPUs, one worker process per GPU,\n'
                        'longest contig first; the output is the same text in the same order')
    p.add_argument('--verbosity', metavar='LEVEL', default='WARNING',
                   help='Set logging level (default: %(default)s)\nUse `INFO` to show work progress')
    p.add_argument('--version', action='version', version='%(prog)s ' + __version__)
    return p


def process(argv=None):
    args = get_argparser().parse_args(argv)
    logger.setLevel(getattr(logging, args.verbosity.upper()))
    logger.info("Pasio:" + str(args))
    splitter = configure_splitter(**vars(args))
    split_bedgraph(in_filename=args.bedgraph, out_filename=args.output_file, splitter=splitter,
                   split_at_gaps=args.split_at_gaps, output_mode=args.output_mode, devices=args.devices)


def main():
    try:
        process()
    except KeyboardInterrupt:
        logger.error('Program was interrupted')
        sys.exit(1)
