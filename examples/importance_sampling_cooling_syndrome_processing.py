#!/usr/bin/env python
"""The reference's importance_sampling_cooling_syndrome_processing.py (Dress_cooling with sympathetic cooling between the
timesteps, locations [12, 18, 24, 78, 24, 119], :8-34) on the B200 backend: the same stages as
examples/importance_sampling_syndrome_processing.py, with the cooling model's parameters.

    python examples/importance_sampling_cooling_syndrome_processing.py [--runs 1000] [--cutoff 1e-6] [--out imp_samp]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from importance_sampling_syndrome_processing import run_pipeline          # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=1000)                     # the reference's `runs` (:7)
    ap.add_argument("--cutoff", type=float, default=1e-6)
    ap.add_argument("--out", default="imp_samp")
    a = ap.parse_args()
    run_pipeline("Dress_cooling", a.runs, a.cutoff, a.out)
