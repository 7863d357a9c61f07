"""Empty stand-in: render.py imports imageio at module top (image IO only)."""
