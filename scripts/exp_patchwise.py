"""GPU experiment: per-kernel-class split of the patchwise path (gather + predict_proba on a 100 000-voxel batch)."""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
g = torch.Generator(device="cuda").manual_seed(5)
shape = (192,) * 3
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)
idx = torch.arange(n, device="cuda") + 40 * 192 * 192
xyz = torch.stack([idx // (192 * 192), (idx // 192) % 192, idx % 192], 1).to(torch.int32).contiguous()
for _ in range(2):
    out = ctx.forward_from_volume(vol, atlas, xyz)
torch.cuda.synchronize()
ctx.set_option("profile", 1)
ctx.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = ctx.forward_from_volume(vol, atlas, xyz)
e1.record()
torch.cuda.synchronize()
print("ms per batch %.2f  -> %.3f M voxels/s" % (e0.elapsed_time(e1) / 3, n / (e0.elapsed_time(e1) / 3) / 1e3))
for k, (ms, c) in sorted(ctx.profile_read().items(), key=lambda kv: -kv[1][0]):
    print("  %-14s %8.2f ms  %5d launches" % (k, ms / 3, c // 3))
