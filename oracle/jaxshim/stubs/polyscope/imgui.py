"""Empty stand-in for polyscope.imgui (GUI only)."""
