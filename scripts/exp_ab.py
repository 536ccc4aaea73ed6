"""GPU A/B experiment inside ONE process (power capping makes run-to-run comparisons unreliable): alternate library
option settings over the same 256^3 dense pass and report the mean time of each.
usage: python scripts/exp_ab.py name1:opt=val,opt=val name2:opt=val ... [--size N] [--rounds R]"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets
args = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and not sys.argv[i - 1].startswith("--")]
size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 256
rounds = int(sys.argv[sys.argv.index("--rounds") + 1]) if "--rounds" in sys.argv else 4
mask_frac = float(sys.argv[sys.argv.index("--mask") + 1]) if "--mask" in sys.argv else 0.0   # candidate mask: centred ball with this volume fraction
configs = []
for a in args:
    name, _, rest = a.partition(":")
    configs.append((name, [(kv.split("=")[0], int(kv.split("=")[1])) for kv in rest.split(",") if kv]))
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(5)
shape = (size,) * 3
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)
lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
proba = torch.zeros(shape + (15,), dtype=torch.float32, device="cuda") if "--proba" in sys.argv else None   # out_probabilities=True
mask = None
if mask_frac > 0:
    ax = torch.arange(size, device="cuda", dtype=torch.float32) - (size - 1) / 2
    r2 = ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2
    rad = size * (3 * mask_frac / (4 * 3.14159265)) ** (1 / 3)
    mask = (r2 < rad * rad).to(torch.uint8).contiguous()
    print("mask: %.3f of the voxels" % float(mask.float().mean()))
ref = None
tot = {n: [] for n, _ in configs}
for r in range(rounds + 1):
    for name, opts in configs:
        for k, v in opts:
            ctx.set_option(k, v)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.segment_volume(vol, atlas, cand_mask=mask, label_vol=lab, proba_vol=proba)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            tot[name].append(e0.elapsed_time(e1))
        if ref is None:
            ref = lab.clone()
        elif r == 0:
            print(name, "labels differing from the first config:", int((lab != ref).sum()))
for name, _ in configs:
    t = tot[name]
    print("%-16s mean %.2f ms  min %.2f  (%s)" % (name, sum(t) / len(t), min(t), " ".join("%.1f" % x for x in t)))
