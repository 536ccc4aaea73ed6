#!/usr/bin/env python
"""bench_train.py -- BASELINE.json config 4/5a: one training step (patch_size 32, global batch 256 by default),
data-parallel over N GPUs with the gradient all-reduce over NCCL.  Secondary to bench.py (the headline
inference metric); same launch convention (torchrun for N > 1), one JSON line on rank 0.

    python bench_train.py [--gpus N] [--global-batch 256] [--steps 20] [--warmup 5]
"""
import argparse
import json
import os

# stdout carries exactly ONE JSON line: everything libraries print (NCCL's version banner ...) is rerouted to stderr
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit_json(obj):
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())

import pickle
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sub-cortical_segmentation_b200")]
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cnn_cort import _native, nets, parallel
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _native.Context(local)
    with open(os.path.join(ROOT, "nets", "miccai2012_v1", "miccai2012_v1.pkl"), "rb") as f:
        ctx.load_weights(nets.pack_params(pickle.load(f, encoding="latin1")))
    n = args.global_batch // world
    g = torch.Generator(device="cuda").manual_seed(rank)
    x = [torch.randn((n, 1, 32, 32), device="cuda", generator=g) for _ in range(3)]
    at = torch.softmax(3 * torch.randn((n, 15), device="cuda", generator=g), 1)
    y = torch.randint(0, 15, (n,), device="cuda", generator=g, dtype=torch.uint8)
    hx = [t.cpu().pin_memory() for t in x] + [at.cpu().pin_memory(), y.cpu().pin_memory()]
    grads = ctx.grad_tensor()
    loss = torch.zeros(1, device="cuda")

    def step(i, from_host=False):
        d = [t.cuda(non_blocking=True) for t in hx] if from_host else x + [at, y]
        ctx.train_forward_backward(*d, n_global=args.global_batch, seed=i, loss_out=loss)
        parallel.allreduce_gradients(grads, loss)
        ctx.adam_step(lr=1e-3, stat_scale=1.0 / world)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.counter("launches")
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    launches = ctx.counter("launches") - l0
    e0.record()
    for i in range(args.steps):
        step(i, from_host=True)
        float(loss.item())
    e1.record()
    sync()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    ar_ms = None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        sync()
        e0.record()
        for _ in range(50):
            dist.all_reduce(grads)
        e1.record()
        sync()
        ar_ms = e0.elapsed_time(e1) / 50
    if rank == 0:
        per = float(ms) / args.steps
        emit_json({"metric": "training_samples_per_sec", "value": args.global_batch / (per * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "training step, patch_size=32, global batch %d, DP over %d GPU(s), per-GPU BN statistics" %
                                                 (args.global_batch, world), "per_gpu_batch": n},
                          "e2e": {"value": args.global_batch / (float(ms2) / args.steps * 1e-3), "unit": "samples/s",
                                  "h2d_bytes_per_step": int(n * (3 * 4096 + 60 + 1)), "d2h_bytes_per_step": 4},
                          "allreduce_ms": ar_ms, "allreduce_bytes": 883455 * 4, "gpu_launches": int(launches), "loss": float(loss.item()),
                          "flops_per_sample_fwd_bwd": 3 * 35407800})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
