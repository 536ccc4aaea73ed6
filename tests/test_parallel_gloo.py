"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding arithmetic and the gradient
all-reduce plumbing (the kernels themselves need a GPU; see tests/test_gpu_*.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cnn_cort import parallel


def test_shard_arithmetic():
    for n in (0, 1, 7, 64, 100003):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                a, b = parallel.shard_range(n, r, world)
                assert 0 <= a <= b <= n and (b - a) - n // world in (0, 1)
                cover += list(range(a, b)) if n < 1000 else [a, b]
            if n < 1000:
                assert cover == list(range(n))
    vols = ["v%02d" % i for i in range(64)]
    got = [parallel.shard_items(vols, r, 8) for r in range(8)]
    assert all(len(g) == 8 for g in got) and sorted(sum(got, [])) == vols
    box = (10, 33, 0, 40, 5, 9)
    slabs = [parallel.shard_box(box, r, 4) for r in range(4)]
    assert slabs[0][0] == 10 and slabs[-1][1] == 33 and all(s[2:] == box[2:] for s in slabs)
    assert all(slabs[i][1] == slabs[i + 1][0] for i in range(3))
    assert parallel.shard_box((0, 2, 0, 1, 0, 1), 3, 4) is None
    idx = np.arange(256)
    parts = [parallel.shard_batch(idx, r, 8) for r in range(8)]
    assert all(len(p) == 32 for p in parts) and sorted(np.concatenate(parts)) == list(idx)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.dist_info()[:2] == (rank, world)
        # a fake flat gradient buffer: each rank holds the gradient of its shard, already divided by the GLOBAL batch
        rng = np.random.RandomState(0)
        per_sample = rng.randn(8, 1000).astype(np.float32)             # 8 samples of one global minibatch
        mine = parallel.shard_batch(np.arange(8), rank, world)
        grads = torch.from_numpy(per_sample[mine].sum(0) / 8.0)
        loss = torch.tensor([float(len(mine)) / 8.0])
        w = parallel.allreduce_gradients(grads, loss)
        assert w == world
        assert torch.allclose(grads, torch.from_numpy(per_sample.mean(0)), atol=1e-6)
        assert abs(float(loss) - 1.0) < 1e-6
        # sharded inference bookkeeping: the slabs of all ranks tile the box exactly once
        slab = parallel.shard_box((3, 20, 0, 5, 0, 5), rank, world)
        cnt = torch.zeros(20)
        cnt[slab[0]:slab[1]] += 1
        dist.all_reduce(cnt)
        assert cnt[3:20].eq(1).all() and cnt[:3].eq(0).all()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
