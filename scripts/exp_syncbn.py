import os, pickle, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np, torch
import torch.distributed as dist
from cnn_cort import _native, nets
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29544", RANK="0", WORLD_SIZE="1")
dist.init_process_group("gloo", rank=0, world_size=1)
n = 24
ctx = _native.Context(0)
with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
    ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
rng = np.random.RandomState(5)
x = [torch.from_numpy(rng.randn(n, 1, 32, 32).astype(np.float32)).cuda() for _ in range(3)]
at = torch.from_numpy(rng.dirichlet(np.ones(15) * 0.3, size=n).astype(np.float32)).cuda()
y = torch.from_numpy(rng.randint(0, 15, n).astype(np.uint8)).cuda()
masks = torch.from_numpy((rng.rand(n, 2700) < 0.5).astype(np.uint8)).cuda()
ctx.set_option("train_graph", 0)
print("no hook   ", float(ctx.train_forward_backward(*x, at, y, drop_masks=masks)))
calls = []
def ident(user, ptr, count, stream):
    calls.append((ptr, count, stream))
    return 0
cb = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)(ident)
_native._check(ctx.lib.sc_set_allreduce_hook(ctx.h, ctypes.cast(cb, ctypes.c_void_p), None))
print("identity  ", float(ctx.train_forward_backward(*x, at, y, drop_masks=masks)), len(calls), calls[:3])
def peek(user, ptr, count, stream):
    t = torch.as_tensor(_native._CudaArrayHolder(ptr, int(count), "<f8"), device="cuda:0")
    with torch.cuda.stream(torch.cuda.ExternalStream(stream or 0, device=0)):
        h = t.cpu()
    if len(calls) < 40: calls.append((float(h[0]), float(h[1]), float(h[192])))
    return 0
calls.clear()
cb2 = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)(peek)
_native._check(ctx.lib.sc_set_allreduce_hook(ctx.h, ctypes.cast(cb2, ctypes.c_void_p), None))
print("peek      ", float(ctx.train_forward_backward(*x, at, y, drop_masks=masks)), calls[:6])
ctx.set_sync_bn(True)
print("sync w=1  ", float(ctx.train_forward_backward(*x, at, y, drop_masks=masks)))
