"""f-4: connected-component post-processing on the device (sc_post_process) vs the scipy restatement of base.py:460-480,
bit for bit, including ties between components, absent classes and classes without overlap (the reference's argmax == 0 quirk)."""
import numpy as np
import pytest
import torch

from oracle import postprocess as op
from gpu_util import cuda_ctx, dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cuda_ctx()
    yield c
    c.close()


def _blobs(shape, rng, n_blobs, classes):
    seg = np.zeros(shape, np.uint8)
    for _ in range(n_blobs):
        c = [rng.randint(0, s) for s in shape]
        r = rng.randint(1, 6)
        sl = tuple(slice(max(0, ci - r), min(s, ci + r + 1)) for ci, s in zip(c, shape))
        seg[sl] = classes[rng.randint(len(classes))]
    return seg


@pytest.mark.parametrize("shape,classes,seed", [((24, 20, 18), list(range(1, 15)), 0), ((40, 36, 30), [1, 2, 3, 5, 8, 13, 14], 1),
                                                ((33, 17, 9), [4], 2), ((64, 64, 64), list(range(1, 15)), 3)])
def test_post_process_equals_scipy(ctx, shape, classes, seed):
    rng = np.random.RandomState(seed)
    seg = _blobs(shape, rng, 60, classes)
    noise = rng.rand(*shape) < 0.03                              # isolated voxels: many single-voxel components, ties
    seg[noise] = np.array(classes)[rng.randint(len(classes), size=int(noise.sum()))]
    mask = np.zeros(shape, np.float32)
    mask[shape[0] // 4:3 * shape[0] // 4, shape[1] // 4:3 * shape[1] // 4, :] = 1.0
    want = op.post_process_segmentation(mask, seg)
    got = ctx.post_process(dev(seg), dev((mask != 0).view(np.uint8))).cpu().numpy()
    assert np.array_equal(got, want)


def test_quirks_absent_class_and_no_overlap(ctx):
    """class 14 absent and class 3 entirely outside the mask: both select `labels == 0`"""
    shape = (16, 16, 16)
    seg = np.zeros(shape, np.uint8)
    seg[2:5, 2:5, 2:5] = 3            # outside the mask
    seg[8:12, 8:12, 8:12] = 7         # inside
    seg[13, 13, 13] = 7               # a second, smaller component of class 7 inside the mask
    mask = np.zeros(shape, np.float32)
    mask[7:, 7:, 7:] = 1
    want = op.post_process_segmentation(mask, seg)
    got = ctx.post_process(dev(seg), dev((mask != 0).view(np.uint8))).cpu().numpy()
    assert np.array_equal(got, want)
    assert (want == 14).sum() > 0 and got[13, 13, 13] == 14       # the small component is dropped and painted over


def test_post_process_segmentation_api_and_full_size(ctx, tmp_path):
    from cnn_cort import base, nifti
    rng = np.random.RandomState(5)
    shape = (256, 256, 256)
    seg = _blobs(shape, rng, 400, list(range(1, 15))).astype(np.float32)      # test_scan hands over a T1-typed label image
    mask = np.zeros(shape, np.float32)
    mask[64:192, 64:192, 64:192] = 1
    d = tmp_path / "s01" / "tmp"
    d.mkdir(parents=True)
    nifti.Nifti1Image(mask, np.eye(4)).to_filename(str(d / "MNI_subcortical_mask.nii.gz"))
    got = base.post_process_segmentation(str(tmp_path / "s01"), seg, device=0)
    want = op.post_process_segmentation(mask, seg)
    assert got.dtype == seg.dtype and np.array_equal(got, want)
