"""f-2: scan preparation on the device (csrc/prep.cu) vs numpy, bit for bit: array-order import of NIfTI (Fortran-ordered)
arrays, the normalisation of base.py:358 (numpy's dtype promotion and pairwise summation order), candidate mask,
bounding box."""
import numpy as np
import pytest
import torch

from gpu_util import cuda_ctx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = cuda_ctx()
    yield c
    c.close()


def _np_norm(image):
    nz = image[np.nonzero(image)]
    return ((image - nz.mean()) / nz.std()).astype(np.float32), nz.mean(), nz.std()


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("shape,channels", [((5, 4, 3), 1), ((33, 17, 9), 1), ((40, 35, 70), 15), ((64, 64, 64), 1)])
def test_import_volume_is_a_bit_exact_reorder(ctx, dtype, shape, channels):
    rng = np.random.RandomState(len(shape) + channels)
    full = shape + ((channels,) if channels > 1 else ())
    a = np.asfortranarray((rng.rand(*full) * 200 - 20).astype(dtype))        # what nifti.load / nibabel hold in memory
    raw, dt = ctx.upload_volume(a, channels=channels)
    got = raw.cpu().numpy().view(dtype).reshape(full)
    assert dt == a.dtype and np.array_equal(got, a)
    raw2, _ = ctx.upload_volume(np.ascontiguousarray(a), channels=channels)   # C-ordered input: plain upload
    assert torch.equal(raw, raw2)


@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("dtype,shape,channels,box", [
    (np.float32, (40, 35, 70), 15, (3, 29, 0, 35, 11, 64)),        # box touching two faces
    (np.float32, (40, 35, 70), 15, (0, 40, 0, 35, 0, 70)),         # the whole volume
    (np.float32, (33, 17, 9), 15, (32, 33, 16, 17, 8, 9)),         # one voxel in the far corner
    (np.int16, (64, 48, 40), 1, (5, 50, 7, 41, 2, 39)),
    (np.float64, (20, 24, 28), 3, (1, 19, 2, 3, 0, 28)),
])
def test_upload_volume_box_writes_exactly_the_box(ctx, order, dtype, shape, channels, box):
    """crop mode: only the candidates' bounding box of the (page-locked or pageable) host priors is uploaded"""
    rng = np.random.RandomState(sum(box))
    full = shape + ((channels,) if channels > 1 else ())
    a = (rng.rand(*full) * 200 - 20).astype(dtype)
    a = np.asfortranarray(a) if order == "F" else np.ascontiguousarray(a)
    tdt = torch.from_numpy(np.zeros(1, dtype)).dtype
    out = torch.full(full, 7, dtype=tdt, device="cuda")
    got = ctx.upload_volume_box(a, box, channels=channels, out=out)
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    ref = np.full(full, 7, dtype)
    sl = (slice(box[0], box[1]), slice(box[2], box[3]), slice(box[4], box[5]))
    ref[sl] = a[sl]
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint8, np.int32])
@pytest.mark.parametrize("shape,zero_frac", [((3, 2, 1), 0.0), ((4, 4, 4), 0.5), ((16, 9, 11), 0.3), ((37, 41, 29), 0.2),
                                             ((128, 96, 80), 0.45), ((64, 64, 64), 0.0)])
def test_normalise_matches_numpy_bit_for_bit(ctx, dtype, shape, zero_frac):
    rng = np.random.RandomState(sum(shape) + int(zero_frac * 10))
    hi = 250 if dtype == np.uint8 else 1500
    img = (rng.rand(*shape) * hi + (0.5 if np.dtype(dtype).kind == "f" else 1)).astype(dtype)
    img[rng.rand(*shape) < zero_frac] = 0
    if np.dtype(dtype).kind == "f":
        img.flat[0] = -0.0 if img.size > 8 else img.flat[0]                   # -0.0 is not a non-zero voxel (numpy)
    ref, m, s = _np_norm(img)
    raw, dt = ctx.upload_volume(np.asfortranarray(img))
    vol, mean, std = ctx.normalise_volume(raw, dt, shape)
    assert mean == float(m) and std == float(s), (mean, float(m), std, float(s))
    assert np.array_equal(vol.cpu().numpy(), ref, equal_nan=True)


def test_normalise_full_size_float32_and_int16(ctx):
    """BASELINE size: 256^3 (16.7 M non-zero values: a 131 072-leaf pairwise tree)."""
    from cnn_cort import synthetic
    t1 = synthetic.make_t1((256, 256, 256), 1234)
    for img in (t1, np.rint(t1).astype(np.int16)):
        ref, m, s = _np_norm(img)
        raw, dt = ctx.upload_volume(img)
        vol, mean, std = ctx.normalise_volume(raw, dt, img.shape)
        assert mean == float(m) and std == float(s)
        assert np.array_equal(vol.cpu().numpy(), ref)
        del vol, raw
    torch.cuda.empty_cache()


def test_all_zero_volume_gives_nan_like_numpy(ctx):
    img = np.zeros((6, 5, 4), np.float32)
    raw, dt = ctx.upload_volume(img)
    vol, mean, std = ctx.normalise_volume(raw, dt, img.shape)
    assert np.isnan(mean) and np.isnan(std) and np.isnan(vol.cpu().numpy()).all()


@pytest.mark.parametrize("dtype", [np.float32, np.int16, np.uint8])
def test_candidate_mask_and_bbox(ctx, dtype):
    rng = np.random.RandomState(9)
    shape = (50, 44, 38)
    img = np.zeros(shape, dtype)
    img[7:31, 3:40, 11:12] = (rng.rand(24, 37, 1) * 100 + 1).astype(dtype)
    img[30, 39, 37] = 5
    raw, dt = ctx.upload_volume(img)
    mask = ctx.candidate_mask(raw, dt, shape)
    assert np.array_equal(mask.cpu().numpy().astype(bool), img.astype(bool))
    box, n = ctx.mask_bbox(mask)
    nz = np.nonzero(img)
    assert n == len(nz[0])
    assert box == (nz[0].min(), nz[0].max() + 1, nz[1].min(), nz[1].max() + 1, nz[2].min(), nz[2].max() + 1)
    empty = torch.zeros(shape, dtype=torch.uint8, device="cuda")
    assert ctx.mask_bbox(empty) == (None, 0)
