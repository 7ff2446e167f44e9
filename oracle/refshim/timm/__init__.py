"""Minimal stand-in for timm 0.4.12, ONLY so that oracle/make_golden.py can import the reference's
models/vig.py and models/TGCN.py in this container (timm is not installed).  Not product code."""
