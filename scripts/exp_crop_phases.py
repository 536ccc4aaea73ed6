"""GPU experiment: phase breakdown of the crop-mode test_scan hot section (synchronising between the phases)."""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
import bench
from cnn_cort import _native, nets, base, synthetic
ctx = _native.Context(0)
with open(bench.WEIGHTS, "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
t1, norm, atlas = bench.synthetic_volume(256, 1234)
def pinned_f(a):
    t = torch.empty(a.size * a.itemsize, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(a.dtype).reshape(a.shape, order="F")
    v[...] = a
    return t, v
k1, t1f = pinned_f(t1); k2, atf = pinned_f(atlas); k3, mkf = pinned_f(synthetic.make_mask(atlas))
shape = t1.shape
def lap(name, t0):
    torch.cuda.synchronize(); t = time.perf_counter()
    print("  %-34s %.2f ms" % (name, (t - t0) * 1e3)); return t
for rep in range(2):
    print("rep", rep)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    raw, dt = ctx.upload_volume(t1f); t0 = lap("T1 upload + reorder", t0)
    vol, mean, std = ctx.normalise_volume(raw, dt, shape); t0 = lap("normalise", t0)
    mraw, mdt = ctx.upload_volume(mkf); t0 = lap("mask upload + reorder", t0)
    cm = ctx.candidate_mask(mraw, mdt, shape); t0 = lap("mask != 0", t0)
    cand = ctx.dilate_mask(cm, 10); t0 = lap("dilate x 10", t0)
    box, n = ctx.mask_bbox(cand); t0 = lap("bbox", t0)
    d_atlas = ctx.upload_volume_box(atf, box, channels=15); t0 = lap("atlas box upload + reorder", t0)
    lab = torch.zeros(shape, dtype=torch.uint8, device="cuda"); t0 = lap("zeros", t0)
    ctx.set_option("profile", 1); ctx.profile_read()
    ctx.segment_volume(vol, d_atlas, box=box, cand_mask=cand, label_vol=lab); t0 = lap("segment_volume (box, %d cand)" % n, t0)
    print("    classes:", {k: round(v[0], 2) for k, v in ctx.profile_read().items()}); ctx.set_option("profile", 0)
    h = base._pinned_out('lab', shape, torch.uint8); h.copy_(lab, non_blocking=True); t0 = lap("label download", t0)
