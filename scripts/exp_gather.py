"""GPU A/B experiment for the gather kernel: option settings alternated in one process, median of many calls.
usage: python scripts/exp_gather.py name:opt=val ... [--n 100000]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native
n = int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 100000
cfgs = []
for i, a in enumerate(sys.argv[1:], 1):
    if a.startswith("--") or sys.argv[i - 1].startswith("--"):
        continue
    name, _, rest = a.partition(":")
    cfgs.append((name, [(kv.split("=")[0], int(kv.split("=")[1])) for kv in rest.split(",") if kv]))
ctx = _native.Context(0)
g = torch.Generator(device="cuda").manual_seed(1)
shape = (256,) * 3
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g)
idx = torch.arange(n, device="cuda") + 100 * 256 * 256
xyz = torch.stack([idx // 65536, (idx // 256) % 256, idx % 256], 1).to(torch.int32).contiguous()
res = {c[0]: [] for c in cfgs}
for rnd in range(6):
    for name, opts in cfgs:
        for k, v in opts:
            ctx.set_option(k, v)
        out = ctx.gather_patches(vol, xyz, atlas=atlas, bg_fix=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ctx.gather_patches(vol, xyz, atlas=atlas, bg_fix=True)
        e1.record()
        torch.cuda.synchronize()
        if rnd:
            res[name].append(e0.elapsed_time(e1) / 5)
        del out
for name, _ in cfgs:
    t = sorted(res[name])
    med = t[len(t) // 2]
    print("%-12s median %.4f ms  best %.4f  -> %.0f GB/s (median), %.0f (best)" % (name, med, t[0], n * 12424 / med / 1e6, n * 12424 / t[0] / 1e6))
