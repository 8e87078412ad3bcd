"""Shared bits of the example harnesses: puts the repository root on sys.path.  The examples run on the CUDA engine
only (no CPU path); tests/test_examples.py drives their `main(argv)` on a GPU-less box by monkeypatching the engine."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BACKEND = "B200 CUDA engine"
