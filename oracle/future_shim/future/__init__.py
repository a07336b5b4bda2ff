"""Stub of the `future` package (not installed in this image) so the unmodified
reference at /root/reference/src can be imported by oracle/make_golden.py.
Test infrastructure only."""
