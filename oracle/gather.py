"""Oracle: candidate-voxel indexing, orthogonal patch gather, atlas vectors.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  numpy restatement of the
reference's data layer; every function names the reference lines it follows
(paths relative to ``/root/reference``).
"""
import random as _random

import numpy as np

PATCH = 32
VIEWS = ("axial", "coronal", "saggital")  # the reference's spelling, base.py:291-296


def normalise(image, dtype=None):
    """(image - mean_nz) / std_nz over the non-zero voxels.

    Test path ``cnn_cort/base.py:358`` applies numpy's own promotion to whatever
    dtype the NIfTI holds (integer T1 -> float64, float32 T1 -> float32); the
    patches are cast to float32 afterwards (:383-385), which is bit-identical to
    casting the normalised volume once.  Train path ``base.py:146`` casts the
    image to float32 first: pass ``dtype=np.float32``.  Under the reference's numpy
    1.12.1 (requirements.txt:18) a float32 ARRAY combined with float64 SCALARS stays
    float32 (value-based casting: the scalars are rounded to float32), whereas an
    integer array with a float64 scalar promotes to float64 -- so the train path is
    float32 arithmetic and the test path on an integer T1 is float64 arithmetic.
    NumPy 2 (NEP 50) would promote the train path to float64; the casts below
    restore the reference's result.
    """
    nz = image[np.nonzero(image)]
    if dtype is not None:
        dt = np.dtype(dtype).type
        return (image.astype(dtype) - dt(nz.mean())) / dt(nz.std())
    return (image - nz.mean()) / nz.std()


def get_mask_voxels(mask, size=None, rng=None):
    """C-ordered coordinates of the non-zero voxels -> int64 [N, 3].

    ``cnn_cort/base.py:310-331``: ``np.stack(np.nonzero(mask), axis=1)`` (x
    slowest, z fastest); with ``size`` the list is shuffled (``random.shuffle``,
    unseeded in the reference, :327-329) and truncated.  ``rng`` (a
    ``random.Random``) makes the shuffle reproducible for tests.
    """
    idx = np.stack(np.nonzero(mask), axis=1).astype(np.int64)
    if size is not None:
        order = list(range(idx.shape[0]))
        (rng or _random).shuffle(order)
        idx = idx[order[:size]]
    return idx


def _view_axes(mode):
    # base.py:291-296: axial -> (p, p, 1), coronal -> (p, 1, p), saggital -> (1, p, p)
    if mode == "axial":
        return (0, 1), 2
    if mode == "coronal":
        return (0, 2), 1
    if mode == "saggital":
        return (1, 2), 0
    raise ValueError("unknown view %r" % (mode,))


def get_patches_loop(image, centers, patch_size=(PATCH, PATCH), mode="axial"):
    """Per-centre slicing exactly as ``cnn_cort/base.py:272-308`` does it.

    Window on an in-plane axis: [c - p//2, c + p - p//2); single index on the
    third axis; the volume is zero-padded by (p//2, p - p//2) first (:298-303),
    so everything outside the volume reads 0.  Slow; small inputs only.
    """
    (a0, a1), a2 = _view_axes(mode)
    full = [1, 1, 1]
    full[a0], full[a1] = patch_size[0], patch_size[1]
    half = [s // 2 for s in full]
    padded = np.pad(image, [(h, s - h) for h, s in zip(half, full)], mode="constant")
    out = []
    for c in centers:
        sl = tuple(slice(int(ci), int(ci) + s) for ci, s in zip(c, full))  # == (c+h)-h .. (c+h)+(s-h)
        out.append(np.squeeze(padded[sl]))
    return out


def get_patches(image, centers, patch_size=(PATCH, PATCH), mode="axial"):
    """Vectorised equivalent of :func:`get_patches_loop` -> [N, p0, p1] (image dtype)."""
    centers = np.asarray(centers, dtype=np.int64).reshape(-1, 3)
    (a0, a1), a2 = _view_axes(mode)
    p0, p1 = patch_size
    h0, h1 = p0 // 2, p1 // 2
    pad = [(0, 0)] * 3
    pad[a0], pad[a1] = (h0, p0 - h0), (h1, p1 - h1)
    padded = np.pad(image, pad, mode="constant")
    i = np.arange(p0)[None, :, None]
    j = np.arange(p1)[None, None, :]
    idx = [None, None, None]
    idx[a0] = centers[:, a0][:, None, None] + i
    idx[a1] = centers[:, a1][:, None, None] + j
    idx[a2] = centers[:, a2][:, None, None] + 0 * i
    return padded[tuple(idx)]


def atlas_vectors_test(atlas, centers):
    """``atlas[x, y, z, :]`` as float32 with the background fix.

    ``cnn_cort/base.py:387-394``: rows whose 15 priors sum to exactly 0 get
    channel 14 set to 1 (inference path only).
    """
    c = np.asarray(centers, dtype=np.int64).reshape(-1, 3)
    v = atlas[c[:, 0], c[:, 1], c[:, 2]].astype(np.float32)
    for r in range(v.shape[0]):
        if np.sum(v[r]) == 0:
            v[r, 14] = 1
    return v


def atlas_vectors_train(atlas, centers):
    """``cnn_cort/base.py:208-218``: plain lookup.  The background fix there sums a
    whole subject's array and indexes an undefined name, so it never applies
    (SURVEY.md quirk Q4)."""
    c = np.asarray(centers, dtype=np.int64).reshape(-1, 3)
    return atlas[c[:, 0], c[:, 1], c[:, 2]]


def patch_batches(image_norm, atlas, centers, batch_size, patch_size=(PATCH, PATCH), datatype=np.float32):
    """The body of ``load_patch_batch`` after I/O (``cnn_cort/base.py:379-397``):
    yields ``(axial, coronal, saggital, atlas_vec, centers)`` with patches cast to
    ``datatype`` and stacked to [n, 1, p, p]."""
    centers = np.asarray(centers, dtype=np.int64).reshape(-1, 3)
    for i in range(0, centers.shape[0], batch_size):
        c = centers[i:i + batch_size]
        views = [get_patches(image_norm, c, patch_size, m).astype(datatype)[:, None] for m in VIEWS]
        yield views[0], views[1], views[2], atlas_vectors_test(atlas, c), c


def candidates(image, crop_mask=None):
    """Candidate voxels of the inference path, ``cnn_cort/base.py:367-372``:
    ``binary_dilation(mask, iterations=10)`` of the registered sub-cortical mask when
    cropping, else the non-zero voxels of the raw T1."""
    if crop_mask is not None:
        from scipy import ndimage
        return get_mask_voxels(ndimage.binary_dilation(crop_mask, iterations=10).astype(bool))
    return get_mask_voxels(image.astype(bool))


def training_vectors(image_norm, labels, rng=None, patch_size=(PATCH, PATCH), balance_neg=True):
    """One subject of ``load_patch_vectors`` (``cnn_cort/base.py:153-181``):
    every voxel with label 1..14, then as many label-15 voxels (shuffled, truncated);
    T1 and label patches for the three views; positives first, negatives after."""
    pos = get_mask_voxels(np.logical_and(labels > 0, labels < 15))
    neg = get_mask_voxels(labels == 15, size=len(pos) if balance_neg else None, rng=rng)
    cen = np.concatenate([pos, neg])
    x = {m: get_patches(image_norm, cen, patch_size, m) for m in VIEWS}
    y = {m: get_patches(labels, cen, patch_size, m) for m in VIEWS}
    return x, y, cen


def generate_training_set(x_axial, x_coronal, x_saggital, x_atlas, y, randomize=True, seed=None):
    """``cnn_cort/base.py:77-110``: concatenate subjects, centre label of the axial label
    patch, 15 -> 0, one seed re-applied to five permutations, add the channel axis."""
    xa = np.concatenate(x_axial, axis=0).astype("float32")
    xc = np.concatenate(x_coronal, axis=0).astype("float32")
    xs = np.concatenate(x_saggital, axis=0).astype("float32")
    xt = np.concatenate(x_atlas, axis=0).astype("float32")
    yt = np.concatenate(y, axis=0).astype("uint8")
    yt = np.squeeze(yt[:, yt.shape[1] // 2, yt.shape[2] // 2])
    yt[yt == 15] = 0
    if randomize:
        if seed is None:
            seed = np.random.randint(np.iinfo(np.int32).max)
        out = []
        for arr in (xa, xc, xs, yt, xt):
            np.random.seed(seed)
            out.append(np.random.permutation(arr))
        xa, xc, xs, yt, xt = out
    return xa[:, None], xc[:, None], xs[:, None], xt, yt
