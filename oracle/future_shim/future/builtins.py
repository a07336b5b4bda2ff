from builtins import *  # noqa: F401,F403
range = range
