"""GPU experiment: where do the roles of the persistent tcgen05 kernel wait?  (per-CTA clock64 counters of the LAST launch
of the chosen kernel class; classes: conv2=5 conv3=6 conv4=7 conv5=8 gemm_d1=9 gemm_fc1=10 gemm_fc2=11)"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
from cnn_cort import _native, nets
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(5)
shape = (size,) * 3
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)
lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
ctx.segment_volume(vol, atlas, label_vol=lab)
names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "epi_wait_tfull", "epi_total", "tiles"]
for cls, nm in [(5, "conv2"), (6, "conv3"), (9, "gemm_d1"), (10, "gemm_fc1"), (11, "gemm_fc2")]:
    ctx.set_option("tc_timing", cls)
    ctx.segment_volume(vol, atlas, label_vol=lab)
    torch.cuda.synchronize()
    v = np.array([[ctx.counter("tc_timing:%d" % (c * 8 + k)) for k in range(8)] for c in range(0, 74, 18)], dtype=np.float64)
    m = v.mean(0)
    print("%-9s tiles/CTA %5.0f | per tile: total %6.0f  mma waits full %6.0f tempty %6.0f | producer waits empty %6.0f | epilogue waits tfull %6.0f (busy %6.0f)" % (
        nm, m[7], m[4] / m[7], m[2] / m[7], m[3] / m[7], m[0] / m[7], m[5] / m[7], (m[6] - m[5]) / m[7]), flush=True)
