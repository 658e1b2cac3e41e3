"""Stub of matplotlib for the drop-in tests (the reference imports pyplot at module level, function.py:12)."""
