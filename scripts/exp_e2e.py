"""GPU experiment: device-resident sc_segment_volume vs host-buffer sc_segment_volume_host (pinned), alternated in one process."""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
from cnn_cort import _native, nets
size = 256
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(5)
shape = (size,) * 3
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)
mask = torch.ones(shape, dtype=torch.uint8, device="cuda")
lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
h_vol, h_atlas, h_mask = vol.cpu().pin_memory(), atlas.cpu().pin_memory(), mask.cpu().pin_memory()
h_lab = torch.zeros(shape, dtype=torch.uint8).pin_memory()
res = {"device": [], "host": [], "host_wall": [], "pageable_wall": []}
p_vol, p_atlas, p_mask = h_vol.numpy().copy(), h_atlas.numpy().copy(), h_mask.numpy().copy()   # plain (pageable) numpy arrays
p_lab = np.zeros(shape, np.uint8)
for r in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.segment_volume(vol, atlas, cand_mask=mask, label_vol=lab); e1.record(); torch.cuda.synchronize()
    if r: res["device"].append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    e0.record(); ctx.segment_volume_host(h_vol.numpy(), h_atlas.numpy(), cand_mask=h_mask.numpy(), label_out=h_lab.numpy()); e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if r: res["host"].append(e0.elapsed_time(e1)); res["host_wall"].append((t1 - t0) * 1e3)
    t0 = time.perf_counter()
    ctx.segment_volume_host(p_vol, p_atlas, cand_mask=p_mask, label_out=p_lab)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    if r: res["pageable_wall"].append((t1 - t0) * 1e3)
    assert np.array_equal(p_lab, h_lab.numpy())
for k, v in res.items():
    print("%-10s mean %.2f ms (%s)" % (k, sum(v) / len(v), " ".join("%.1f" % x for x in v)))
