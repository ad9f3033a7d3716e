"""Launcher for the full end-point parity run (tests/endpoint_parity.py; it executes the oracle, so the code lives under tests/):
    python profiles/tools/endpoint_parity.py --scenes c1,c2,c3 --iters 3000 --out profiles/r02_endpoint_parity.json"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv[0] = os.path.join(ROOT, "tests", "endpoint_parity.py")
runpy.run_path(sys.argv[0], run_name="__main__")
