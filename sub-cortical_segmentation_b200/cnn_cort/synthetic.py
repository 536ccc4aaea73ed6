"""Synthetic subjects for benchmarks and tests: a strictly positive smooth
T1, a 15-channel atlas prior made of Gaussian blobs (with all-zero rows far from every
blob, to exercise the background fix), the registered-mask file the crop path reads and
a 15-class label volume for training.  Not part of the reference; it only pre-creates the
files whose existence makes the reference skip registration (base.py:201,362).
"""
import os

import numpy as np

from . import nifti


def make_t1(shape, seed=1234):
    from scipy.ndimage import gaussian_filter
    rng = np.random.RandomState(seed)
    f = gaussian_filter(rng.standard_normal(shape).astype(np.float32), 3.0)
    f /= f.std()
    f += 0.15 * rng.standard_normal(shape).astype(np.float32)
    return np.maximum(f * 150.0 + 600.0, 1.0).astype(np.float32)


def make_atlas(shape, seed=1234):
    """-> (atlas [X,Y,Z,15] float32, blob centres [14,3], sigmas [14])"""
    rng = np.random.RandomState(seed + 7)
    X, Y, Z = shape
    atlas = np.zeros(shape + (15,), np.float32)
    box = np.array([min(96, X), min(96, Y), min(64, Z)]) // 2
    mid = np.array(shape) // 2
    cen = np.stack([rng.randint(mid[a] - box[a], mid[a] + box[a] + 1, size=14) for a in range(3)], 1)
    sig = rng.uniform(6.0, 12.0, size=14) * min(1.0, min(shape) / 128.0 + 0.25)
    near = np.zeros(shape, bool)
    for j in range(14):
        g = [np.exp(-0.5 * ((np.arange(shape[a]) - cen[j, a]) / sig[j]) ** 2).astype(np.float32) for a in range(3)]
        lo = [max(0, int(cen[j, a] - 4 * sig[j])) for a in range(3)]
        hi = [min(shape[a], int(cen[j, a] + 4 * sig[j]) + 1) for a in range(3)]
        sl = tuple(slice(l, h) for l, h in zip(lo, hi))
        blob = g[0][sl[0], None, None] * g[1][None, sl[1], None] * g[2][None, None, sl[2]]
        blob[blob < 1e-3] = 0
        atlas[sl + (j,)] = blob
        r = int(40 * min(1.0, min(shape) / 128.0 + 0.25))
        nl = [max(0, cen[j, a] - r) for a in range(3)]
        nh = [min(shape[a], cen[j, a] + r + 1) for a in range(3)]
        near[tuple(slice(l, h) for l, h in zip(nl, nh))] = True
    s = atlas[..., :14].sum(-1)
    big = s > 1.0
    atlas[big, :14] /= s[big, None]
    atlas[..., 14] = np.clip(1.0 - atlas[..., :14].sum(-1), 0.0, 1.0)
    atlas[~near] = 0.0
    return atlas, cen, sig


def make_mask(atlas):
    from scipy.ndimage import binary_dilation
    m = atlas[..., 0:13].sum(-1) > 0.01
    return binary_dilation(m, iterations=5).astype(np.float32)


def make_labels(atlas):
    """labels 1..14 where a structure prior > 0.5, a 2-voxel ring of 15 around them, else 0"""
    from scipy.ndimage import binary_dilation
    best = atlas[..., :14].argmax(-1)
    val = atlas[..., :14].max(-1)
    lab = np.where(val > 0.5, best + 1, 0).astype(np.uint8)
    ring = binary_dilation(lab > 0, iterations=2) & (lab == 0)
    lab[ring] = 15
    return lab


def write_subject(root, name, shape=(256, 256, 256), seed=1234, with_labels=False, t1_name="T1.nii.gz",
                  roi_name="gt_15_classes.nii.gz", zoom=1.0, t1_dtype=np.float32):
    d = os.path.join(root, name)
    os.makedirs(os.path.join(d, "tmp"), exist_ok=True)
    aff = np.diag([zoom, zoom, zoom, 1.0])
    t1 = make_t1(shape, seed)
    if np.dtype(t1_dtype).kind in "iu":       # scanner-style integer intensities (the usual on-disk type of a T1)
        t1 = np.rint(t1).astype(t1_dtype)
    atlas, _, _ = make_atlas(shape, seed)
    nifti.Nifti1Image(t1, aff).to_filename(os.path.join(d, t1_name))
    nifti.Nifti1Image(atlas, aff).to_filename(os.path.join(d, "tmp", "MNI_sub_probabilities.nii.gz"))
    nifti.Nifti1Image(make_mask(atlas), aff).to_filename(os.path.join(d, "tmp", "MNI_subcortical_mask.nii.gz"))
    if with_labels:
        nifti.Nifti1Image(make_labels(atlas), aff).to_filename(os.path.join(d, roi_name))
    return d
