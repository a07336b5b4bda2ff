"""Window geometry over a candidate array (reference: /root/reference/src/pasio/dto/sliding_window.py:4-15)."""
from __future__ import division


class SlidingWindow(object):
    def __init__(self, window_size, window_shift):
        self.window_size = window_size
        self.window_shift = window_shift

    def ranges(self, length):
        """[start, stop) index ranges: starts 0, shift, 2*shift, ... < length-1; at most size+1 long."""
        return [(start, min(start + self.window_size + 1, length))
                for start in range(0, length - 1, self.window_shift)]

    def windows(self, arr):
        length = len(arr)
        for start, stop in self.ranges(length):
            yield (arr[start:stop], stop / length)
