"""Minimal stand-in for NVIDIA Apex so the UNMODIFIED reference sources can run offline.

Apex is not installed in this image and cannot be fetched (the reference's Makefile clones it from
GitHub).  Only ``apex.amp`` is provided, with Apex O1 semantics rebuilt on torch.amp:
per-call fp16 casts (autocast), fp32 master weights, dynamic loss scaling with Apex's constants
(2**16 initial, x2 per 2000 clean steps, /2 and skip the step on overflow).
This file is benchmark plumbing, never imported by the product package.
"""
from . import amp  # noqa: F401
