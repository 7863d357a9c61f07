"""jax.scipy stand-in: import placeholder only (render.py imports it; hot path never calls it)."""
