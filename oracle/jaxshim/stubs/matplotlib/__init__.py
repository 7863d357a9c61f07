"""Empty stand-in: queries.py imports matplotlib.pyplot at module top (unused)."""
