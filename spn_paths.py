"""Puts the repo root on sys.path (bench.py and the tools import the package and oracle/ from a source checkout)."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
