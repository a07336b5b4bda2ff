"""Logger `pasio` with round / window context prefixes.

Same observable behaviour as the reference's /root/reference/src/pasio/logging.py:4-30: messages
are prefixed with `Round N:` and `Window [a, b):` when those context keys are set, written to
stderr, default level WARNING.  Per-window lines only exist on the generic (host-callback)
route; the fused device route reports per-round candidate counts.
"""
from __future__ import absolute_import
import logging


class LoggingContextFilter(object):
    def __init__(self):
        self.context = {}

    def put_to_context(self, key, value):
        self.context[key] = value

    def remove_from_context(self, key):
        self.context.pop(key, None)

    def filter(self, record):
        prefix = ''
        if 'round' in self.context:
            prefix += 'Round %d: ' % self.context['round']
        if 'window' in self.context:
            prefix += 'Window %s: ' % self.context['window']
        if prefix:
            record.msg = prefix + str(record.msg)
        return True


logger = logging.getLogger('pasio')
if not logger.handlers:
    _handler = logging.StreamHandler()
    _handler.setFormatter(logging.Formatter('%(asctime)s - %(levelname)s - %(message)s'))
    logger.addHandler(_handler)
    logger.setLevel(logging.WARNING)

logging_filter = LoggingContextFilter()
logger.addFilter(logging_filter)
