"""GPU experiment: CTA-pair (cta_group::2) tcgen05 variant vs the persistent single-CTA kernel on the head layers."""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
from cnn_cort import _native, nets
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
x = torch.randn((n, 576), device="cuda", generator=g)
x[:, 555:] = 0
for which in (3, 4, 0):
    ctx.set_option("tc_variant", 2)
    ref = ctx.dense_layer(which, x, 1)
    torch.cuda.synchronize()
    for variant in (2, 3):
        ctx.set_option("tc_variant", variant)
        out = ctx.dense_layer(which, x, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            out = ctx.dense_layer(which, x, 1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5 * 1e3
        print("layer %d variant %d: %.3f ms  max|diff vs variant 2| %.3e" % (which, variant, dt, float((out - ref).abs().max())), flush=True)
