"""Empty stand-in: the reference's utils.py imports polyscope at module top (GUI only)."""
