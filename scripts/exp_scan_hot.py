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

# phase breakdown of the full-brain call (synchronising between the phases)
import time
from cnn_cort import base
def pinned_f(a):
    t = torch.empty(a.size * a.itemsize, dtype=torch.uint8, pin_memory=True)
    v = t.numpy().view(a.dtype).reshape(a.shape, order="F")
    v[...] = a
    return t, v
k1, t1f = pinned_f(t1)
k2, atf = pinned_f(atlas)
shape = t1.shape
def lap(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    print("  %-28s %.2f ms" % (name, (t - t0) * 1e3))
    return t
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    raw, dt = ctx.upload_volume(t1f); t0 = lap("T1 upload + reorder", t0)
    d_atlas, _ = ctx.upload_volume(atf, channels=15); t0 = lap("atlas upload + reorder", t0)
    vol, mean, std = ctx.normalise_volume(raw, dt, shape); t0 = lap("normalise", t0)
    cand = ctx.candidate_mask(raw, dt, shape); t0 = lap("candidate mask", t0)
    box, n = ctx.mask_bbox(cand); t0 = lap("bbox", t0)
    lab = torch.zeros(shape, dtype=torch.uint8, device="cuda"); t0 = lap("zeros", t0)
    ctx.segment_volume(vol, d_atlas.view(torch.float32).view(shape + (15,)), box=box, cand_mask=cand, label_vol=lab); t0 = lap("segment_volume", t0)
    h = base._pinned_out('lab', shape, torch.uint8); h.copy_(lab, non_blocking=True); t0 = lap("label download", t0)

def run(name, n=4):
    tm = {}
    base.segment_arrays(ctx, t1f, atf, None, False, tm)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        base.segment_arrays(ctx, t1f, atf, None, False, tm)
    print("%-40s %.1f ms" % (name, (time.perf_counter() - t0) / n * 1e3), flush=True)
run("segment_arrays (side-stream atlas)")
orig = base._side_stream
base._side_stream = lambda dev: torch.cuda.current_stream()
run("segment_arrays (atlas on the main stream)")
base._side_stream = orig
run("segment_arrays (side-stream atlas)")
# device-resident pass alone, same process
vol = torch.from_numpy(norm).cuda(); da = torch.from_numpy(atlas).cuda(); lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
for _ in range(2): ctx.segment_volume(vol, da, label_vol=lab)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): ctx.segment_volume(vol, da, label_vol=lab)
torch.cuda.synchronize(); print("sc_segment_volume device-resident %.1f ms" % ((time.perf_counter() - t0) / 4 * 1e3))

ctx.set_option("profile", 1)
for name, fn in (("side", orig), ("main", lambda dev: torch.cuda.current_stream()), ("side", orig)):
    base._side_stream = fn
    tm = {}
    base.segment_arrays(ctx, t1f, atf, None, False, tm)
    ctx.profile_read()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    base.segment_arrays(ctx, t1f, atf, None, False, tm)
    wall = (time.perf_counter() - t0) * 1e3
    prof = ctx.profile_read()
    print("%s wall %.1f | sum %.1f | " % (name, wall, sum(v[0] for v in prof.values())) + "  ".join("%s %.2f" % (k, v[0]) for k, v in prof.items()), flush=True)
base._side_stream = orig
