"""Lazy grouping of consecutive items, cut where condition(prev, cur) holds
(Ruby's Enumerable#slice_when; reference: /root/reference/src/pasio/utils/slice_when.py:3-39).
Like itertools.groupby, a group must be consumed before the next one is requested; an
unfinished group is drained automatically."""


class slice_when(object):
    _END = object()

    def __init__(self, iterable, condition):
        self._it = iter(iterable)
        self._condition = condition
        self._pending = slice_when._END   # first item of the next group
        self._started = False
        self._group = None

    def __iter__(self):
        return self

    def __next__(self):
        if not self._started:
            self._started = True
            self._pending = next(self._it, slice_when._END)
        elif self._group is not None:
            for _ in self._group:
                pass
        if self._pending is slice_when._END:
            raise StopIteration
        self._group = self._emit_group()
        return self._group

    next = __next__

    def _emit_group(self):
        current = self._pending
        self._pending = slice_when._END
        while True:
            yield current
            following = next(self._it, slice_when._END)
            if following is slice_when._END:
                return
            if self._condition(current, following):
                self._pending = following
                return
            current = following
