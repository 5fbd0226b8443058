"""Stand-in for the absent third-party package torchOptics (reference env.py:24-25).

Only the five symbols the hot path uses are provided, with the semantics restated in
DESIGN.md section 5 (parity with the real package is unpinned: its source is not available).
"""
from . import optics, metrics  # noqa: F401
