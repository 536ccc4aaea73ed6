"""The BASELINE.json configurations themselves against the oracle:

  configs[1]  the bench's synthetic 256^3 T1 + atlas priors (cnn_cort.synthetic, seed 1234), every voxel a candidate:
              probabilities at >= 2 000 sampled voxels (corners, faces, all-zero prior rows, blob interiors) vs the fp64
              oracle, through sc_segment_volume and sc_segment_volume_host
  configs[4]  a 0.7 mm 320^3 volume with out_probabilities=True (1.97 GB probability volume)
  configs[0]  test_scan with speedup_segmentation=True on a 256^3 synthetic subject vs oracle candidates + oracle labels
  a-4         load_patch_batch yields == oracle.gather.patch_batches, bit for bit

Tolerance (north_star): softmax within 1e-3 absolute, argmax agreement >= 99.9 %; indices and patches bit-exact."""
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


def _synthetic(size, seed=1234):
    """bench.py's workload: strictly positive smooth T1 normalised like base.py:358, blob atlas with all-zero rows"""
    from cnn_cort import synthetic
    shape = (size,) * 3
    t1 = synthetic.make_t1(shape, seed)
    norm = og.normalise(t1).astype(np.float32)
    atlas, cen, sig = synthetic.make_atlas(shape, seed)
    return t1, norm, atlas, cen


def _sample_voxels(shape, atlas, blob_centres, n, rng):
    X, Y, Z = shape
    pts = [(x, y, z) for x in (0, X - 1) for y in (0, Y - 1) for z in (0, Z - 1)]                 # corners
    pts += [(0, Y // 2, Z // 2), (X - 1, Y // 2, Z // 2), (X // 2, 0, Z // 2), (X // 2, Y - 1, Z // 2),
            (X // 2, Y // 2, 0), (X // 2, Y // 2, Z - 1), (X - 1, Y - 1, Z - 2), (15, 16, 17), (X - 16, Y - 17, Z - 15)]
    zero_rows = np.stack(np.nonzero(atlas.sum(-1) == 0), 1)
    pts += [tuple(r) for r in zero_rows[rng.choice(len(zero_rows), size=min(200, len(zero_rows)), replace=False)]]
    for c in blob_centres:                                                                           # unsaturated priors
        off = rng.randint(-14, 15, size=(40, 3))
        pts += [tuple(np.clip(c + o, 0, np.array(shape) - 1)) for o in off]
    rest = max(0, n - len(pts))
    pts += [tuple(r) for r in np.stack([rng.randint(0, s, size=rest) for s in shape], 1)]
    return np.unique(np.asarray(pts, dtype=np.int64), axis=0)


def _oracle_proba(P, norm, atlas, pick):
    x = [og.get_patches(norm, pick, (32, 32), m).astype(np.float32)[:, None] for m in og.VIEWS]
    return on.forward(P, *x, og.atlas_vectors_test(atlas, pick), dtype=torch.float64)


def _check(got, ref, what):
    err = np.abs(got - ref).max()
    agree = (np.argmax(got, 1) == np.argmax(ref, 1)).mean()
    assert err < TOL, "%s: max|dp| = %g" % (what, err)
    assert agree >= 0.999, "%s: argmax agreement %g" % (what, agree)
    return err


def test_config2_bench_volume_256_vs_oracle(ctx, P):
    t1, norm, atlas, cen = _synthetic(256)
    shape = t1.shape
    rng = np.random.RandomState(2)
    pick = _sample_voxels(shape, atlas, cen, 2100, rng)
    assert len(pick) >= 2000 and (atlas[pick[:, 0], pick[:, 1], pick[:, 2]].sum(-1) == 0).sum() >= 100
    ref = _oracle_proba(P, norm, atlas, pick)
    d_vol, d_atlas = dev(norm), dev(atlas)
    d_mask = dev((t1 != 0).view(np.uint8))
    assert int(d_mask.sum()) == t1.size                                      # every voxel is a candidate (quirk Q3)
    lab = torch.full(shape, 99, dtype=torch.uint8, device="cuda")
    prob = torch.zeros(shape + (15,), dtype=torch.float32, device="cuda")
    ctx.segment_volume(d_vol, d_atlas, cand_mask=d_mask, label_vol=lab, proba_vol=prob)
    got = prob[pick[:, 0], pick[:, 1], pick[:, 2]].cpu().numpy()
    _check(got, ref, "256^3 device-resident")
    L = lab.cpu().numpy()
    assert L.max() <= 14
    assert np.array_equal(L[pick[:, 0], pick[:, 1], pick[:, 2]], np.argmax(got, 1))
    # the labels are the arg-max of the probabilities everywhere (first maximum wins), checked on the device
    assert bool((prob.view(-1, 15).argmax(1).to(torch.uint8) == lab.view(-1)).all())
    del prob, d_vol, d_atlas, d_mask
    torch.cuda.empty_cache()
    # the host-buffer entry point (what bench.py's e2e number times): identical labels, same probabilities
    lab_h, prob_h = ctx.segment_volume_host(norm, atlas, cand_mask=(t1 != 0).view(np.uint8), want_proba=True)
    assert np.array_equal(lab_h, L)
    _check(prob_h[pick[:, 0], pick[:, 1], pick[:, 2]], ref, "256^3 host entry point")
    torch.cuda.empty_cache()


def test_config5_volume_320_with_probabilities_vs_oracle(ctx, P):
    """0.7 mm 320^3 volume, out_probabilities=True: 32.8 M voxels, a 1.97 GB probability volume (byte offsets up to 2^31)."""
    t1, norm, atlas, cen = _synthetic(320, seed=77)
    shape = t1.shape
    rng = np.random.RandomState(5)
    pick = _sample_voxels(shape, atlas, cen, 700, rng)
    far = np.stack([rng.randint(300, 320, size=60), rng.randint(0, 320, size=60), rng.randint(0, 320, size=60)], 1)   # the last x-planes
    pick = np.unique(np.concatenate([pick, far]), axis=0)
    ref = _oracle_proba(P, norm, atlas, pick)
    lab = torch.full(shape, 99, dtype=torch.uint8, device="cuda")
    prob = torch.full(shape + (15,), -1.0, dtype=torch.float32, device="cuda")
    ctx.segment_volume(dev(norm), dev(atlas), label_vol=lab, proba_vol=prob)
    got = prob[pick[:, 0], pick[:, 1], pick[:, 2]].cpu().numpy()
    _check(got, ref, "320^3 with probabilities")
    assert np.array_equal(lab[pick[:, 0], pick[:, 1], pick[:, 2]].cpu().numpy(), np.argmax(got, 1))
    assert int(lab.max()) <= 14 and float(prob.min()) >= 0.0                # every voxel written
    s = prob.view(-1, 15)[:: 4099].sum(1)
    assert float((s - 1).abs().max()) < 1e-5
    del lab, prob
    torch.cuda.empty_cache()


def _options(root, **kw):
    o = {'experiment': 'miccai2012_v1', 'patch_size': [32, 32], 'mode': 'cuda0', 'device': 0, 'load_weights': 'True',
         'net_verbose': 0, 'train_split': 0.25, 'max_epochs': 1, 'patience': 1, 'batch_size': 128,
         'test_batch_size': 100000, 'debug': 'False', 'out_probabilities': 'False', 'post_process': 'False',
         'crop': 'True', 'crop_bool': True, 'test_folder': root, 't1_name': 'T1.nii.gz'}
    o.update(kw)
    return o


def test_config1_test_scan_crop_256_vs_oracle(ctx, P, tmp_path, weights_path):
    """configs[0]: speedup_segmentation=True on a 256^3 subject through the drop-in test_scan: the candidates are the
    registered mask dilated 10 times (base.py:367-370), nothing else is written, labels match the oracle's."""
    from cnn_cort import base, nets, nifti, synthetic
    root = str(tmp_path)
    d = synthetic.write_subject(root, "s01", shape=(256, 256, 256), seed=1234)
    options = _options(root, timings={})
    net = nets.build_model(os.path.dirname(os.path.dirname(weights_path)), options)
    t1_names, _ = base.load_test_names(options)
    base.test_scan(net, t1_names[0], options)
    seg = nifti.load(os.path.join(d, 'out_subcortical_rawseg.nii.gz')).get_data()
    t1 = nifti.load(os.path.join(d, 'T1.nii.gz')).get_data()
    atlas = nifti.load(os.path.join(d, 'tmp', 'MNI_sub_probabilities.nii.gz')).get_data()
    mask = nifti.load(os.path.join(d, 'tmp', 'MNI_subcortical_mask.nii.gz')).get_data()
    cen = og.candidates(t1, crop_mask=mask)                                # scipy binary_dilation x 10 + np.nonzero
    assert options['timings']['n_candidates'] == len(cen) and 100000 < len(cen) < 4000000
    cand = np.zeros(t1.shape, bool)
    cand[cen[:, 0], cen[:, 1], cen[:, 2]] = True
    assert (seg[~cand] == 0).all()                                          # nothing written outside the candidates
    lo, hi = cen.min(0), cen.max(0) + 1
    assert tuple(options['timings']['box']) == (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
    rng = np.random.RandomState(8)
    pick = cen[rng.choice(len(cen), size=1500, replace=False)]
    norm = og.normalise(t1)
    assert float(options['timings']['mean_nz']) == float(t1[np.nonzero(t1)].mean())
    ref = _oracle_proba(P, norm, atlas, pick)
    agree = (seg[pick[:, 0], pick[:, 1], pick[:, 2]] == np.argmax(ref, 1)).mean()
    assert agree >= 0.999, agree
    assert seg.dtype == t1.dtype and seg.max() > 0
    # the candidate list the patchwise flow iterates over is the oracle's, in order (device dilation + compaction)
    got_cen = base.get_mask_voxels(base.candidate_mask(t1, d, options), as_array=True)
    assert np.array_equal(got_cen, cen)


@pytest.mark.parametrize("crop,t1_dtype", [(True, np.float32), (False, np.float32), (False, np.int16)])
def test_load_patch_batch_equals_oracle_batches(ctx, tmp_path, crop, t1_dtype):
    """a-4: the generator's yields == the oracle's restatement of base.py:357-397, bit for bit (normalisation included)."""
    from cnn_cort import base, nifti, synthetic
    root = str(tmp_path)
    d = synthetic.write_subject(root, "s01", shape=(44, 40, 36), seed=9, t1_dtype=t1_dtype)
    options = _options(root, crop_bool=crop, test_batch_size=7000)
    t1 = nifti.load(os.path.join(d, 'T1.nii.gz')).get_data()
    if not crop:
        t1 = t1.copy()
        t1[5:9, :, 7] = 0                                                   # non-crop candidates = non-zero T1 voxels
        nifti.Nifti1Image(t1, np.eye(4)).to_filename(os.path.join(d, 'T1.nii.gz'))
    atlas = nifti.load(os.path.join(d, 'tmp', 'MNI_sub_probabilities.nii.gz')).get_data()
    mask = nifti.load(os.path.join(d, 'tmp', 'MNI_subcortical_mask.nii.gz')).get_data()
    cen = og.candidates(t1, crop_mask=mask if crop else None)
    want = list(og.patch_batches(og.normalise(t1), atlas, cen, 7000))
    got = list(base.load_patch_batch(os.path.join(d, 'T1.nii.gz'), options))
    assert len(got) == len(want) >= 2
    for g, w in zip(got, want):
        for k in range(3):
            assert g[k].dtype == np.float32 and g[k].shape == w[k].shape and np.array_equal(g[k], w[k])
        assert np.array_equal(g[3], w[3])
        assert np.array_equal(np.asarray(g[4]), w[4]) and isinstance(g[4][0], tuple)
