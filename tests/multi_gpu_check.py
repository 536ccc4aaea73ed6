"""Run under torchrun (one rank per GPU): checks of the N>1 paths on real GPUs.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
from cnn_cort import _native, nets, parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _native.Context(local)
    with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
        ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))

    # 1. inference sharded by x-slab: union of the ranks' outputs == unsharded result, no collective on the data path
    g = torch.Generator(device="cuda").manual_seed(7)
    shape = (48, 40, 36)
    vol = torch.randn(shape, device="cuda", generator=g)
    atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
    atlas = atlas / atlas.sum(-1, keepdim=True)
    full = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    ctx.segment_volume(vol, atlas, label_vol=full)
    part = torch.full(shape, 255, dtype=torch.uint8, device="cuda")
    slab = parallel.segment_volume_sharded(ctx, vol, atlas, label_vol=part)
    assert bool((part[slab[0]:slab[1]] == full[slab[0]:slab[1]]).all())
    touched = (part != 255).to(torch.int32)
    dist.all_reduce(touched)                       # test-only: every voxel written exactly once across ranks
    assert bool((touched == 1).all())

    # 2. data-parallel training: all-reduced gradient == sum of the ranks' shard gradients, parameters stay identical
    rng = np.random.RandomState(3)
    n = 32
    x = [rng.randn(n, 1, 32, 32).astype(np.float32) for _ in range(3)]
    at = rng.dirichlet(np.ones(15) * 0.3, size=n).astype(np.float32)
    y = rng.randint(0, 15, n).astype(np.uint8)
    idx = parallel.shard_batch(np.arange(n), rank, world)
    d = [torch.from_numpy(a[idx]).cuda() for a in x] + [torch.from_numpy(at[idx]).cuda(), torch.from_numpy(y[idx]).cuda()]
    grads = ctx.grad_tensor()
    for step in range(3):
        loss = ctx.train_forward_backward(*d, n_global=n, seed=100 + step)
        local_g = grads.clone()
        parallel.allreduce_gradients(grads, loss)
        gathered = [torch.zeros_like(local_g) for _ in range(world)]
        dist.all_gather(gathered, local_g)
        assert torch.allclose(grads, torch.stack(gathered).sum(0), rtol=1e-5, atol=1e-7)
        ctx.adam_step(lr=1e-3, stat_scale=1.0 / world)
    p = ctx.param_tensor().clone()
    ref = p.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(p, ref), "parameters diverged across ranks"
    assert torch.isfinite(p).all() and np.isfinite(float(loss))
    dist.barrier()
    if rank == 0:
        print("multi_gpu_check ok: world=%d loss=%.4f" % (world, float(loss)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
