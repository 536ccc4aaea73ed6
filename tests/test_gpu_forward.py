"""K2/K3 parity: CUDA forward (patchwise and dense) through the C-ABI vs the fp64 oracle.
Tolerance (BASELINE.json north_star): softmax within 1e-3 absolute, argmax agreement >= 99.9 %."""
import os

import numpy as np
import pytest
import torch

from oracle import gather as og, network as on
from gpu_util import cuda_ctx, dev

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def P(weights_path):
    return on.load_params(weights_path)


@pytest.fixture(scope="module")
def ctx(P):
    from cnn_cort import nets
    c = cuda_ctx()
    c.load_weights(nets.pack_params(P))
    yield c
    c.close()


def _backends(ctx):
    return [0, 1] if ctx.counter("gemm") == 1 else [0]


def _soft_atlas(rng, n):
    return rng.dirichlet(np.ones(15) * 0.3, size=n).astype(np.float32)


def _check(got, ref, what):
    err = np.abs(got - ref).max()
    agree = (np.argmax(got, 1) == np.argmax(ref, 1)).mean()
    assert err < TOL, "%s: max|dp| = %g" % (what, err)
    assert agree >= 0.999, "%s: argmax agreement %g" % (what, agree)


def test_forward_vs_golden_and_oracle(ctx, P, golden_dir):
    G = np.load(os.path.join(golden_dir, "forward_golden.npz"))
    x = [og.get_patches(G["vol"], G["centers"], (32, 32), m).astype(np.float32)[:, None] for m in og.VIEWS]
    default = ctx.counter("gemm")
    for be in _backends(ctx):
        ctx.set_option("gemm", be)
        proba, label = ctx.forward(*[dev(a) for a in x], dev(G["atlas"]))
        _check(proba.cpu().numpy(), G["proba64"], "golden backend %d" % be)
        assert np.array_equal(label.cpu().numpy(), np.argmax(proba.cpu().numpy(), 1))
        rng = np.random.RandomState(11)
        at = _soft_atlas(rng, len(G["centers"]))   # unsaturated softmax: the harsh case
        proba, _ = ctx.forward(*[dev(a) for a in x], dev(at))
        _check(proba.cpu().numpy(), on.forward(P, *x, at, dtype=torch.float64), "soft atlas backend %d" % be)
    ctx.set_option("gemm", default)


@pytest.mark.parametrize("n,kind", [(1, "normal"), (129, "normal"), (200, "zero"), (64, "large")])
def test_forward_edge_inputs(ctx, P, n, kind):
    rng = np.random.RandomState(n)
    scale = {"normal": 1.0, "zero": 0.0, "large": 6.0}[kind]
    x = [(rng.randn(n, 1, 32, 32) * scale).astype(np.float32) for _ in range(3)]
    at = _soft_atlas(rng, n)
    ref = on.forward(P, *x, at, dtype=torch.float64)
    default = ctx.counter("gemm")
    for be in _backends(ctx):
        ctx.set_option("gemm", be)
        proba, _ = ctx.forward(*[dev(a) for a in x], dev(at))
        got = proba.cpu().numpy()
        assert np.isfinite(got).all() and np.allclose(got.sum(1), 1, atol=1e-5)
        _check(got, ref, "%s n=%d backend %d" % (kind, n, be))
    ctx.set_option("gemm", default)


def test_empty_batch_and_one_hot_atlas(ctx):
    z = torch.zeros((0, 1, 32, 32), device="cuda")
    proba, label = ctx.forward(z, z, z, torch.zeros((0, 15), device="cuda"))
    assert proba.shape == (0, 15) and label.shape == (0,)
    rng = np.random.RandomState(3)
    x = [dev((rng.randn(15, 1, 32, 32) * 0.3).astype(np.float32)) for _ in range(3)]
    _, label = ctx.forward(*x, dev(np.eye(15, dtype=np.float32)))
    assert list(label.cpu().numpy()) == list(range(1, 15)) + [0]


def test_host_and_from_volume_entry_points_agree(ctx):
    rng = np.random.RandomState(21)
    vol = rng.randn(30, 28, 26).astype(np.float32)
    atlas = _soft_atlas(rng, vol.size).reshape(vol.shape + (15,))
    atlas[2, 2, 2] = 0
    cen = og.get_mask_voxels(np.ones(vol.shape, bool))[::23].astype(np.int32)
    ax, co, sa, at = ctx.gather_patches(dev(vol), dev(cen), atlas=dev(atlas))
    p_dev, l_dev = ctx.forward(ax, co, sa, at)
    p_host, l_host = ctx.forward_host(ax.cpu().numpy(), co.cpu().numpy(), sa.cpu().numpy(), at.cpu().numpy())
    p_vol, l_vol = ctx.forward_from_volume(dev(vol), dev(atlas), dev(cen))
    assert np.array_equal(p_dev.cpu().numpy(), p_host) and np.array_equal(l_dev.cpu().numpy(), l_host)
    assert np.array_equal(p_dev.cpu().numpy(), p_vol.cpu().numpy()) and np.array_equal(l_dev.cpu().numpy(), l_vol.cpu().numpy())


def test_segment_volume_host_pinned_and_pageable(ctx):
    """sc_segment_volume_host: pinned host buffers (asynchronous chunked atlas upload) and plain numpy arrays (pageable:
    a helper thread issues the blocking copies) give the result of the device-resident call; several slabs, so that the
    per-chunk atlas events are exercised."""
    rng = np.random.RandomState(5)
    shape = (34, 30, 28)
    vol = rng.randn(*shape).astype(np.float32)
    atlas = _soft_atlas(rng, vol.size).reshape(shape + (15,))
    cand = (rng.rand(*shape) < 0.6).view(np.uint8)
    ctx.set_option("chunk_voxels", 5000)
    lab_dev = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    ctx.segment_volume(dev(vol), dev(atlas), cand_mask=dev(cand), label_vol=lab_dev)
    lab_page, _ = ctx.segment_volume_host(vol, atlas, cand_mask=cand)
    pin = [torch.from_numpy(a).pin_memory() for a in (vol, atlas, cand)]
    lab_pin, _ = ctx.segment_volume_host(pin[0].numpy(), pin[1].numpy(), cand_mask=pin[2].numpy())
    ctx.set_option("chunk_voxels", 1 << 20)
    assert np.array_equal(lab_dev.cpu().numpy(), lab_page) and np.array_equal(lab_page, lab_pin)


@pytest.mark.parametrize("shape,box", [((20, 24, 18), None), ((37, 29, 41), (5, 30, 0, 29, 7, 33)), ((16, 40, 12), (3, 4, 10, 11, 5, 6))])
def test_dense_volume_vs_oracle(ctx, P, shape, box):
    rng = np.random.RandomState(sum(shape))
    vol = rng.randn(*shape).astype(np.float32)
    atlas = _soft_atlas(rng, vol.size).reshape(shape + (15,))
    atlas[1, 1, 1] = 0
    cand = rng.rand(*shape) < 0.8
    lab = torch.full(shape, 99, dtype=torch.uint8, device="cuda")
    prob = torch.full(shape + (15,), -1.0, dtype=torch.float32, device="cuda")
    default = ctx.counter("gemm")
    for be in _backends(ctx):
        ctx.set_option("gemm", be)
        lab.fill_(99); prob.fill_(-1.0)
        ctx.segment_volume(dev(vol), dev(atlas), box=box, cand_mask=dev(cand.view(np.uint8)), label_vol=lab, proba_vol=prob)
        inside = np.zeros(shape, bool)
        b = box or (0, shape[0], 0, shape[1], 0, shape[2])
        inside[b[0]:b[1], b[2]:b[3], b[4]:b[5]] = True
        sel = inside & cand
        L, Pv = lab.cpu().numpy(), prob.cpu().numpy()
        assert (L[~sel] == 99).all() and (Pv[~sel] == -1.0).all()      # untouched outside box / mask
        cen = og.get_mask_voxels(sel)
        pick = cen[rng.choice(len(cen), size=min(300, len(cen)), replace=False)]
        x = [og.get_patches(vol, pick, (32, 32), m)[:, None] for m in og.VIEWS]
        ref = on.forward(P, *x, og.atlas_vectors_test(atlas, pick), dtype=torch.float64)
        got = Pv[pick[:, 0], pick[:, 1], pick[:, 2]]
        _check(got, ref, "dense %s backend %d" % (shape, be))
        assert np.array_equal(L[pick[:, 0], pick[:, 1], pick[:, 2]], np.argmax(got, 1))
    ctx.set_option("gemm", default)


def test_dense_backends_agree(ctx):
    """The exact-fp32 SIMT cross-check back-end (sc_set_option("gemm", 0), tests only) and the tcgen05 product path compute
    the same function on a volume large enough for several strips and slabs: probabilities within the tolerance, labels
    equal on >= 99.9 % of the voxels."""
    g = torch.Generator(device="cuda").manual_seed(11)
    shape = (48, 40, 56)
    vol = torch.randn(shape, device="cuda", generator=g)
    atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
    atlas = atlas / atlas.sum(-1, keepdim=True)
    atlas[2, 3, 4] = 0                                       # background fix inside the fused epilogue
    out = []
    for be in (1, 0):
        ctx.set_option("gemm", be)
        prob = torch.zeros(shape + (15,), dtype=torch.float32, device="cuda")
        lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
        ctx.segment_volume(vol, atlas, label_vol=lab, proba_vol=prob)
        out.append((prob, lab))
    ctx.set_option("gemm", 1)
    assert float((out[0][0] - out[1][0]).abs().max()) < TOL
    assert float((out[0][1] == out[1][1]).float().mean()) >= 0.999


@pytest.mark.parametrize("density", [0.02, 0.35, 0.7])
def test_dense_candidate_compaction(ctx, density):
    """With a sparse candidate mask the FC head runs on the compacted candidate rows of every slab; same results as the
    uncompacted pipeline at the candidates, nothing written elsewhere (several slabs, one of them without candidates)."""
    if ctx.counter("gemm") != 1:
        pytest.skip("tcgen05 back-end not selected")
    g = torch.Generator(device="cuda").manual_seed(21)
    shape = (40, 48, 36)
    vol = torch.randn(shape, device="cuda", generator=g)
    atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
    atlas = atlas / atlas.sum(-1, keepdim=True)
    atlas[7, 8, 9] = 0
    cand = (torch.rand(shape, device="cuda", generator=g) < density)
    cand[7, 8, 9] = True
    cand[20:26] = False                                      # whole x-planes (and with small chunks whole slabs) without candidates
    cand = cand.to(torch.uint8).contiguous()
    ctx.set_option("chunk_voxels", 6000)                     # ~3 x-planes per slab
    out = []
    for on in (1, 0):
        ctx.set_option("tc_compact", on)
        prob = torch.full(shape + (15,), -1.0, dtype=torch.float32, device="cuda")
        lab = torch.full(shape, 99, dtype=torch.uint8, device="cuda")
        ctx.segment_volume(vol, atlas, cand_mask=cand, label_vol=lab, proba_vol=prob)
        out.append((prob, lab))
    ctx.set_option("tc_compact", 1)
    ctx.set_option("chunk_voxels", 1 << 20)
    sel = cand.bool()
    assert bool((out[0][1][~sel] == 99).all()) and bool((out[0][0][~sel] == -1.0).all())
    assert float((out[0][0][sel] - out[1][0][sel]).abs().max()) < 1e-5
    assert float((out[0][1][sel] == out[1][1][sel]).float().mean()) >= 0.9999


def test_sparse_mask_sweep_item_skipping(ctx):
    """Sparse candidate masks: the conv sweeps skip the items (strip x row segment) outside the receptive-field reach of every
    candidate.  Candidates in blobs, on faces and in the corners of the volume get bit-identical probabilities with and without the
    skipping, although the workspace still holds the maps of ANOTHER volume in the skipped regions."""
    if ctx.counter("gemm") != 1:
        pytest.skip("tcgen05 back-end not selected")
    g = torch.Generator(device="cuda").manual_seed(33)
    shape = (150, 120, 140)
    other = torch.randn(shape, device="cuda", generator=g) * 5
    vol = torch.randn(shape, device="cuda", generator=g)
    atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
    atlas = atlas / atlas.sum(-1, keepdim=True)
    cand = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    cand[60:75, 40:52, 100:118] = 1                          # a blob
    cand[5:9, 100:110, 3:30] = 1                             # a slab near three faces
    for c in ((0, 0, 0), (149, 119, 139), (0, 119, 0), (149, 0, 139), (75, 0, 70), (0, 60, 139), (149, 60, 0), (33, 77, 99)):
        cand[c] = 1                                          # isolated voxels: corners, faces, interior
    out = []
    for skip in (1, 0):
        lab0 = torch.zeros(shape, dtype=torch.uint8, device="cuda")
        ctx.segment_volume(other, atlas, label_vol=lab0)     # fills the workspace with the maps of a different volume
        ctx.set_option("tc_skip", skip)
        prob = torch.full(shape + (15,), -1.0, dtype=torch.float32, device="cuda")
        lab = torch.full(shape, 99, dtype=torch.uint8, device="cuda")
        ctx.segment_volume(vol, atlas, cand_mask=cand, label_vol=lab, proba_vol=prob)
        out.append((prob, lab))
    ctx.set_option("tc_skip", 1)
    sel = cand.bool()
    assert bool((out[0][1][~sel] == 99).all()) and bool((out[0][0][~sel] == -1.0).all())
    assert torch.equal(out[0][0][sel], out[1][0][sel]) and torch.equal(out[0][1][sel], out[1][1][sel])
    # and against the patchwise path (a different formulation of the same function)
    xyz = torch.nonzero(sel).to(torch.int32)[::7].contiguous()
    pp, _ = ctx.forward_from_volume(vol, atlas, xyz)
    got = out[0][0][xyz[:, 0].long(), xyz[:, 1].long(), xyz[:, 2].long()]
    assert float((got - pp).abs().max()) < 2e-4


def test_dense_equals_patchwise_at_scale(ctx):
    """Size-independent property at a larger size: the dense path and the patchwise path are the
    same function of (volume, atlas, voxel)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    shape = (96, 80, 72)
    vol = torch.randn(shape, device="cuda", generator=g)
    atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 6
    atlas = atlas / atlas.sum(-1, keepdim=True)
    ctx.set_option("chunk_voxels", 100000)   # several head chunks
    prob = torch.zeros(shape + (15,), dtype=torch.float32, device="cuda")
    lab = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    ctx.segment_volume(vol, atlas, label_vol=lab, proba_vol=prob)
    ctx.set_option("chunk_voxels", 1 << 20)
    idx = torch.randint(0, vol.numel(), (4000,), device="cuda", generator=g)
    xyz = torch.stack([idx // (80 * 72), (idx // 72) % 80, idx % 72], 1).to(torch.int32)
    p_patch, l_patch = ctx.forward_from_volume(vol, atlas, xyz)
    p_dense = prob.view(-1, 15)[idx]
    assert float((p_patch - p_dense).abs().max()) < TOL
    assert float((l_patch.long() == lab.view(-1)[idx].long()).float().mean()) >= 0.999


def test_test_scan_end_to_end(ctx, tmp_path, weights_path):
    from cnn_cort import base, nets, nifti, synthetic
    root = str(tmp_path)
    d = synthetic.write_subject(root, "s01", shape=(48, 44, 40), seed=3)
    options = {'experiment': 'miccai2012_v1', 'patch_size': [32, 32], 'mode': 'cuda0', 'device': 0, 'load_weights': 'True',
               'net_verbose': 0, 'train_split': 0.25, 'max_epochs': 1, 'patience': 1, 'batch_size': 128,
               'test_batch_size': 5000, 'debug': 'False', 'out_probabilities': 'True', 'post_process': 'False',
               'crop': 'True', 'crop_bool': True, 'test_folder': root, 't1_name': 'T1.nii.gz'}
    net = nets.build_model(os.path.dirname(os.path.dirname(weights_path)), options)
    t1_names, _ = base.load_test_names(options)
    minutes = base.test_scan(net, t1_names[0], options)
    assert minutes >= 0
    seg_dense = nifti.load(os.path.join(d, 'out_subcortical_rawseg.nii.gz')).get_data()
    prob_dense = nifti.load(os.path.join(d, 'out_subcortical_prob.nii.gz')).get_data()
    options['inference'] = 'patchwise'
    base.test_scan(net, t1_names[0], options)
    seg_patch = nifti.load(os.path.join(d, 'out_subcortical_rawseg.nii.gz')).get_data()
    prob_patch = nifti.load(os.path.join(d, 'out_subcortical_prob.nii.gz')).get_data()
    assert seg_dense.shape == (48, 44, 40) and prob_dense.shape == (48, 44, 40, 15)
    assert (seg_dense == seg_patch).mean() >= 0.999 and np.abs(prob_dense - prob_patch).max() < TOL
    assert seg_dense.max() > 0
    options['post_process'] = 'True'; options['out_probabilities'] = 'False'; options['inference'] = 'dense'
    base.test_scan(net, t1_names[0], options)
    assert os.path.exists(os.path.join(d, 'out_subcortical_seg_prec.nii.gz'))


def test_full_size_volume_properties(ctx):
    """BASELINE size (256^3, every voxel): size-independent properties -- the dense path equals the patchwise path on a
    random sample of voxels, the candidate mask gates the writes, labels are the arg-max of the probabilities."""
    g = torch.Generator(device="cuda").manual_seed(11)
    shape = (256, 256, 256)
    vol = torch.randn(shape, device="cuda", generator=g)
    atlas = torch.rand(shape + (15,), device="cuda", generator=g) ** 8
    atlas = atlas / atlas.sum(-1, keepdim=True)
    mask = (torch.rand(shape, device="cuda", generator=g) < 0.9).to(torch.uint8)
    lab = torch.full(shape, 77, dtype=torch.uint8, device="cuda")
    ctx.segment_volume(vol, atlas, cand_mask=mask, label_vol=lab)
    assert bool((lab[mask == 0] == 77).all()) and int(lab[mask != 0].max()) <= 14
    idx = torch.randint(0, vol.numel(), (3000,), device="cuda", generator=g)
    idx = idx[mask.view(-1)[idx] != 0]
    xyz = torch.stack([idx // 65536, (idx // 256) % 256, idx % 256], 1).to(torch.int32)
    p_patch, l_patch = ctx.forward_from_volume(vol, atlas, xyz)
    agree = float((l_patch.long() == lab.view(-1)[idx].long()).float().mean())
    assert agree >= 0.999, agree
    del vol, atlas, mask, lab
    torch.cuda.empty_cache()
