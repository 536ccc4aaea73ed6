"""GPU experiment: tcgen05 kernel variants (one-tile CTAs vs persistent, kx-reuse descriptor modes) against the
SIMT fp32 path on one dense volume -- max |dp| and time per variant."""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets

size = int(sys.argv[1]) if len(sys.argv) > 1 else 96
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
g = torch.Generator(device="cuda").manual_seed(5)
shape = (size, size - 8, size - 16)
vol = torch.randn(shape, device="cuda", generator=g)
atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
atlas = atlas / atlas.sum(-1, keepdim=True)


def run(**opts):
    for k, v in opts.items():
        ctx.set_option(k, v)
    prob = torch.zeros(shape + (15,), dtype=torch.float32, device="cuda")
    lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    ctx.segment_volume(vol, atlas, label_vol=lab, proba_vol=prob)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.segment_volume(vol, atlas, label_vol=lab, proba_vol=prob)
    torch.cuda.synchronize()
    return prob, lab, (time.perf_counter() - t0) * 1e3


ref, rlab, t = run(gemm=0)
print("simt fp32            %8.1f ms" % t, flush=True)
for name, opts in [("tc one-tile CTAs", dict(gemm=1, tc_variant=1, tc_kx_reuse=0)),
                   ("tc persistent", dict(gemm=1, tc_variant=2, tc_kx_reuse=0)),
                   ("tc persistent kx1", dict(gemm=1, tc_variant=2, tc_kx_reuse=1)),
                   ("tc persistent kx2", dict(gemm=1, tc_variant=2, tc_kx_reuse=2))]:
    p, l, t = run(**opts)
    print("%-20s %8.1f ms  max|dp| %.3e  label agreement %.5f" % (name, t, float((p - ref).abs().max()), float((l == rlab).float().mean())), flush=True)
