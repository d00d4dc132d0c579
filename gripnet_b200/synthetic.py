"""Synthetic supergraph generators, re-exported from the neutral top-level ``synthdata`` module
(kept under this name for the examples and for callers of earlier revisions)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from synthdata import *  # noqa: E402,F401,F403
from synthdata import _mirror, _neg_pairs  # noqa: E402,F401
