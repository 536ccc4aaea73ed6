"""one dense pass over a synthetic cube (for ncu captures): python scripts/one_pass.py [size] [passes]"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(5)
shape = (size,) * 3
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)
lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
for _ in range(passes):
    ctx.segment_volume(vol, atlas, label_vol=lab)
torch.cuda.synchronize()
print("done", int(lab.sum()))
