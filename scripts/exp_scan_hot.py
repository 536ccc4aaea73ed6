"""GPU experiment: the test_scan hot section (bench.py's test_scan_hot) alone: full brain and crop, wall clock per call."""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
import bench
from cnn_cort import _native, nets
ctx = _native.Context(0)
with open(bench.WEIGHTS, "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
t1, norm, atlas = bench.synthetic_volume(256, 1234)
print(bench.bench_test_scan_hot(ctx, torch, t1, atlas, steps=5))
