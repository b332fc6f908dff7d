# Puts the repo root on sys.path for tests/, bench.py and __graft_entry__.py alike.
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
