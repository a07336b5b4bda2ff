"""Bedgraph interval records (reference: /root/reference/src/pasio/dto/intervals.py:5-39)."""
from collections import namedtuple

from ..logging import logger


class ScoredInterval(namedtuple('ScoredInterval', ['start', 'stop', 'mean_count', 'log_marginal_likelyhood'])):
    @property
    def length(self):
        return self.stop - self.start


class BedgraphInterval(namedtuple('BedgraphInterval', ['chrom', 'start', 'stop', 'count'])):
    @property
    def length(self):
        return self.stop - self.start

    @classmethod
    def from_string(cls, line):
        fields = line.split()
        chrom, count_text = fields[0], fields[3]
        try:
            count = int(count_text)
        except ValueError:
            count = int(float(count_text))
            logger.warning("Pasio cannot be used with floating point counts. `%s` was automatically converted "
                           "to an integer `%d` as an approximation. Make sure this value was designed to "
                           "actually be an integer count." % (count_text, count))
        return cls(chrom, int(fields[1]), int(fields[2]), count)

    @classmethod
    def each_in_stream(cls, stream):
        for line in stream:
            line = line.strip()
            if line:
                yield cls.from_string(line)

    @classmethod
    def each_in_file(cls, filename):
        from ..utils.gzip_utils import open_for_read
        with open_for_read(filename) as stream:
            for interval in cls.each_in_stream(stream):
                yield interval
