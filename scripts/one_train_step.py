"""a few training steps (for ncu captures): python scripts/one_train_step.py [n] [steps]"""
import os, pickle, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import torch
from cnn_cort import _native, nets
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
ctx.set_option("train_graph", 0)
g = torch.Generator(device="cuda").manual_seed(0)
x = [torch.randn((n, 1, 32, 32), device="cuda", generator=g) for _ in range(3)]
at = torch.softmax(3 * torch.randn((n, 15), device="cuda", generator=g), 1)
y = torch.randint(0, 15, (n,), device="cuda", generator=g, dtype=torch.uint8)
for i in range(steps):
    ctx.train_forward_backward(*x, at, y, seed=i); ctx.adam_step()
torch.cuda.synchronize()
print("done")
