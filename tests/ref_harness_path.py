"""Puts tests/golden (ref_harness.py, the fixture generators) on sys.path."""
import os
import sys

_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
if _G not in sys.path:
    sys.path.insert(0, _G)
