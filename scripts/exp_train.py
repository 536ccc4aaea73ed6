"""Times the training step (forward + backward + Adam) at a few batch sizes, graph on / off.  python scripts/exp_train.py"""
import os, pickle, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets

ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
for n in [int(a) for a in sys.argv[1:]] or [256, 1024]:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = [torch.randn((n, 1, 32, 32), device="cuda", generator=g) for _ in range(3)]
    at = torch.softmax(3 * torch.randn((n, 15), device="cuda", generator=g), 1)
    y = torch.randint(0, 15, (n,), device="cuda", generator=g, dtype=torch.uint8)
    loss = torch.zeros(1, device="cuda")
    for graph in (0, 1):
        ctx.set_option("train_graph", graph)
        for i in range(5):
            ctx.train_forward_backward(*x, at, y, seed=i, loss_out=loss); ctx.adam_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            ctx.train_forward_backward(*x, at, y, seed=i, loss_out=loss); ctx.adam_step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("n=%d graph=%d: %.3f ms/step  %.0f samples/s  loss %.4f" % (n, graph, ms, n / ms * 1e3, float(loss)))
    ctx.set_option("train_graph", 0); ctx.set_option("profile", 1); ctx.profile_read()
    for i in range(5):
        ctx.train_forward_backward(*x, at, y, seed=i, loss_out=loss); ctx.adam_step()
    print("  per class (5 steps):", {k: (round(v[0] / 5, 3), v[1] // 5) for k, v in ctx.profile_read().items()})
    ctx.set_option("profile", 0)
