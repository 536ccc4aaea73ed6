import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
from cnn_cort import _native, nets
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
rng = np.random.RandomState(5)
x = [torch.from_numpy(rng.randn(n, 1, 32, 32).astype(np.float32)).cuda() for _ in range(3)]
at = torch.from_numpy(rng.dirichlet(np.ones(15) * 0.3, size=n).astype(np.float32)).cuda()
y = torch.from_numpy(rng.randint(0, 15, n).astype(np.uint8)).cuda()
masks = torch.from_numpy((rng.rand(n, 2700) < 0.5).astype(np.uint8)).cuda()
ctx.set_option("train_graph", 0)
try:
    loss = ctx.train_forward_backward(*x, at, y, drop_masks=masks)
    torch.cuda.synchronize()
    print("ok loss", float(loss))
except Exception as e:
    print("FAILED:", str(e)[:600])
