"""The GPU-filling soft-argmax microbenchmark of bench.py on its own (for ncu captures and A/B runs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
peaks, _ = bench.load_peaks()
print(json.dumps(bench.softargmax_roofline(0, peaks)))
