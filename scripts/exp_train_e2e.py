"""GPU experiment: per-step wall clock of the host-fed training step (pinned batch -> device, step, loss read back)."""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
for n in (256, 1024):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = [torch.randn((n, 1, 32, 32), device="cuda", generator=g) for _ in range(3)]
    at = torch.softmax(3 * torch.randn((n, 15), device="cuda", generator=g), 1)
    y = torch.randint(0, 15, (n,), device="cuda", generator=g, dtype=torch.uint8)
    hx = [t.cpu().pin_memory() for t in x] + [at.cpu().pin_memory(), y.cpu().pin_memory()]
    loss = torch.zeros(1, device="cuda")
    for i in range(5):
        ctx.train_forward_backward(*x, at, y, seed=i, loss_out=loss); ctx.adam_step()
    torch.cuda.synchronize()
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        d = [t.cuda(non_blocking=True) for t in hx]
        t1 = time.perf_counter()
        ctx.train_forward_backward(*d, seed=i, loss_out=loss)
        t2 = time.perf_counter()
        ctx.adam_step()
        t3 = time.perf_counter()
        float(loss.item())
        t4 = time.perf_counter()
        ts.append([1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))])
    for r in ts:
        print("n=%d copies %.2f  fwd_bwd call %.2f  adam call %.2f  loss.item %.2f ms" % (n, *r))
