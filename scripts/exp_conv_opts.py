"""GPU experiment: time one dense pass (256^3) under different tcgen05 options; per-kernel-class event times."""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
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
ref = None
for name, opts in [("nacc1", dict(tc_nacc=1)), ("nacc2", dict(tc_nacc=2)), ("nacc4", dict(tc_nacc=4)),
                   ("nacc1 kx0", dict(tc_nacc=1, tc_kx_reuse=0)), ("nacc4 kx0", dict(tc_nacc=4, tc_kx_reuse=0))]:
    ctx.set_option("tc_kx_reuse", 1)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.segment_volume(vol, atlas, label_vol=lab)
    ctx.set_option("profile", 1); ctx.profile_read()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.segment_volume(vol, atlas, label_vol=lab)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) * 1e3
    prof = ctx.profile_read(); ctx.set_option("profile", 0)
    if ref is None:
        ref = lab.clone()
    print("%-10s %7.1f ms  agree %.6f  " % (name, t, float((lab == ref).float().mean())) +
          " ".join("%s=%.1f" % (k, v[0]) for k, v in prof.items() if k.startswith("conv") or k.startswith("gemm")), flush=True)
