"""Training step: tensor-core path (gemm=1) vs the exact-fp32 SIMT cross-check (gemm=0), per gradient tensor.
python scripts/exp_train_parity.py [n]"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
from cnn_cort import _native, nets

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
rng = np.random.RandomState(5)
x = [torch.from_numpy(rng.randn(n, 1, 32, 32).astype(np.float32)).cuda() for _ in range(3)]
at = torch.from_numpy(rng.dirichlet(np.ones(15) * 0.3, size=n).astype(np.float32)).cuda()
y = torch.from_numpy(rng.randint(0, 15, n).astype(np.uint8)).cuda()
masks = torch.from_numpy((rng.rand(n, 2700) < 0.5).astype(np.uint8)).cuda()
out = {}
for be in (0, 1):
    ctx.set_option("gemm", be)
    for graph in ((0, 1) if be == 1 else (0,)):
        ctx.set_option("train_graph", graph)
        loss = ctx.train_forward_backward(*x, at, y, drop_masks=masks)
        torch.cuda.synchronize()
        out[(be, graph)] = (float(loss), nets.unpack_params(ctx.grad_tensor().cpu().numpy()))
ref_loss, ref = out[(0, 0)]
for key in ((1, 0), (1, 1)):
    loss, G = out[key]
    print("backend %d graph %d: loss %.6f (ref %.6f)" % (key[0], key[1], loss, ref_loss))
    worst = 0
    for name, arrs in ref.items():
        for k, a in enumerate(arrs):
            d = np.abs(G[name][k] - a).max() / max(np.abs(a).max(), 1e-6)
            worst = max(worst, d)
            if d > 2e-3 or not np.isfinite(d):
                print("   %-28s[%d] rel err %.3e  |ref|max %.3e" % (name, k, d, np.abs(a).max()))
    print("   worst rel err %.3e" % worst)
