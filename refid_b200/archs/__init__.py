"""Arch plug-in directory: every `*_arch.py` file here is what a BasicSR-style `archs/` folder scan would import."""
