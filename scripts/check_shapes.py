"""GPU check at awkward sizes: the dense path (with and without a sparse candidate mask / bounding box) against the
patchwise path on random voxels.  usage: python scripts/check_shapes.py X Y Z"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (181, 217, 181)
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(3)
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)
X, Y, Z = shape
prob = torch.zeros(shape + (15,), dtype=torch.float32, device="cuda")
lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
ctx.segment_volume(vol, atlas, label_vol=lab, proba_vol=prob)
idx = torch.randint(0, vol.numel(), (20000,), device="cuda", generator=g)
xyz = torch.stack([idx // (Y * Z), (idx // Z) % Y, idx % Z], 1).to(torch.int32).contiguous()
p_patch, l_patch = ctx.forward_from_volume(vol, atlas, xyz)
d = float((p_patch - prob.view(-1, 15)[idx]).abs().max())
agree = float((l_patch.long() == lab.view(-1)[idx].long()).float().mean())
print("full volume %s: max |dp| dense vs patchwise %.2e, label agreement %.5f" % (shape, d, agree))
assert d < 1e-3 and agree >= 0.999
# sparse mask + bounding box
ax = [torch.arange(n, device="cuda", dtype=torch.float32) - (n - 1) / 2 for n in shape]
ball = ((ax[0][:, None, None] / (0.4 * X)) ** 2 + (ax[1][None, :, None] / (0.35 * Y)) ** 2 + (ax[2][None, None, :] / (0.3 * Z)) ** 2 < 1).to(torch.uint8).contiguous()
nz = ball.nonzero()
box = []
for a in range(3):
    box += [int(nz[:, a].min()), int(nz[:, a].max()) + 1]
prob2 = torch.full(shape + (15,), -1.0, dtype=torch.float32, device="cuda")
lab2 = torch.full(shape, 99, dtype=torch.uint8, device="cuda")
ctx.segment_volume(vol, atlas, box=tuple(box), cand_mask=ball, label_vol=lab2, proba_vol=prob2)
sel = ball.bool()
d2 = float((prob2[sel] - prob[sel]).abs().max())
print("masked (%.2f of the voxels, box %s): max |dp| vs full run %.2e, untouched outside: %s" % (
    float(sel.float().mean()), box, d2, bool((lab2[~sel] == 99).all()) and bool((prob2[~sel] == -1).all())))
assert d2 < 1e-4 and bool((lab2[~sel] == 99).all())
print("ok")
