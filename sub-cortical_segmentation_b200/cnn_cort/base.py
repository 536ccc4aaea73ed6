"""Data layer with the reference's signatures (``cnn_cort/base.py``), on the GPU.

Kept names: load_data, load_test_names, generate_training_set, load_patch_vectors,
get_atlas_vectors, load_patches, get_patches, get_mask_voxels, load_patch_batch, test_scan,
post_process_segmentation.  Candidate indexing, patch / atlas gathering, the network and the
result scatter and the 10x dilation of the crop mask run in ``libsubcort_b200.so``; NIfTI
I/O, intensity normalisation and the connected-component post-processing stay on the host
exactly where the reference has them (outside the timed path).

Deliberate deviations from reference quirks: Q1 prediction is not gated
by ``debug``; Q2 ``speedup_segmentation`` is parsed as a boolean; Q10 one forward pass
serves both labels and probabilities; registration (``register_masks``) is not rebuilt --
the ``tmp/MNI_*`` files must exist.
"""
import os
import random
import time

import numpy as np
from scipy import ndimage

from . import _native
from . import nifti as nib
from .nifti import load as load_nii

_contexts = {}


def get_context(device=None):
    """Process-wide sc_ctx per device for the data-layer helpers."""
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _contexts:
        _contexts[device] = _native.Context(device)
    return _contexts[device]


def _dev(a, device, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a if dtype is None else a.astype(dtype, copy=False)))
    return t.to('cuda:%d' % device)


def _centers_tensor(centers, device):
    c = np.asarray(centers, dtype=np.int32).reshape(-1, 3)
    return _dev(c, device)


# ---------------------------------------------------------------------------------------------
# names
# ---------------------------------------------------------------------------------------------
def load_test_names(options):
    """Sorted sub-folders of the inference folder -> (T1 paths, subject names). [base.py:41-50]"""
    dir_name = options['test_folder']
    t1_name = options['t1_name']
    subjects = [f for f in sorted(os.listdir(dir_name)) if os.path.isdir(os.path.join(dir_name, f))]
    t1_names = [os.path.join(dir_name, subject, t1_name) for subject in subjects]
    return t1_names, subjects


# ---------------------------------------------------------------------------------------------
# candidates and patches
# ---------------------------------------------------------------------------------------------
def get_mask_voxels(mask, size=None, device=None, as_array=False):
    """Coordinates of the non-zero voxels in np.nonzero (C) order. [base.py:310-331]

    Returns a list of (x, y, z) tuples like the reference (``as_array=True``: int32 [N,3]).
    ``size``: shuffle (``random.shuffle``) and keep the first ``size`` entries.
    """
    ctx = get_context(device)
    m = np.asarray(mask)
    if m.dtype == np.bool_:
        m = m.view(np.uint8)
    elif m.dtype not in (np.uint8, np.float32, np.int32):
        m = (m != 0).view(np.uint8)
    xyz = ctx.nonzero_coords(_dev(m, ctx.device)).cpu().numpy()
    if size is not None:
        order = list(range(xyz.shape[0]))
        random.shuffle(order)
        xyz = xyz[order[:size]]
    if as_array:
        return xyz
    return [tuple(int(v) for v in r) for r in xyz]


def get_patches(image, centers, patch_size=(32, 32), mode='axial', device=None):
    """32x32 patches of one view around each centre, zeros outside the volume. [base.py:272-308]

    axial = [dx, dy] at z, coronal = [dx, dz] at y, saggital = [dy, dz] at x.
    Returns float32 [N, 32, 32] (the reference returns a list of N [32,32] arrays of the
    image dtype and every caller casts to float32; label volumes pass through exactly
    because labels <= 15 are exact in float32).
    """
    if tuple(patch_size) != (32, 32):
        raise ValueError("only patch_size (32, 32) is implemented")
    ctx = get_context(device)
    views = {'axial': (True, False, False), 'coronal': (False, True, False), 'saggital': (False, False, True)}[mode]
    vol = _dev(np.asarray(image), ctx.device, np.float32)
    out = ctx.gather_patches(vol, _centers_tensor(centers, ctx.device), views=views)
    return out[views.index(True)][:, 0].cpu().numpy()


def load_patch_batch(scan_name, options, datatype=np.float32):
    """Generator over test batches: (axial, coronal, saggital, atlas, centers). [base.py:335-397]

    Arrays are float32 [n,1,32,32] / [n,15] numpy like the reference's.  The volume and the
    atlas are uploaded once and stay on the device for the whole scan.
    """
    import torch
    ctx = get_context(options.get('device'))
    dir_name, name = os.path.split(scan_name)
    image = load_nii(scan_name).get_data()
    atlas_name = os.path.join(dir_name, 'tmp', 'MNI_sub_probabilities.nii.gz')
    _require_registered(atlas_name)
    shape = tuple(int(s) for s in image.shape[:3])
    raw, dt = ctx.upload_volume(image)
    vol, _, _ = ctx.normalise_volume(raw, dt, shape)                       # base.py:358, numpy's result bit for bit
    if _crop_enabled(options):                                             # base.py:367-372
        mraw, mdt = ctx.upload_volume(nib.load(os.path.join(dir_name, 'tmp', 'MNI_subcortical_mask.nii.gz')).get_data())
        cand = ctx.dilate_mask(ctx.candidate_mask(mraw, mdt, shape), 10)
    else:
        cand = ctx.candidate_mask(raw, dt, shape)
    centers = ctx.nonzero_coords(cand)                                     # np.nonzero order, stays on the device
    if options['debug'] == 'True':
        print("    -->  num of samples to test:", len(centers))
    atlas = np.asarray(load_nii(atlas_name).get_data())
    if atlas.dtype != np.float32:
        atlas = atlas.astype(np.float32, order='K')
    d_atlas, _ = ctx.upload_volume(atlas, channels=15)
    d_atlas = d_atlas.view(torch.float32).view(shape + (15,))
    batch_size = options['test_batch_size']
    for i in range(0, len(centers), batch_size):
        c = centers[i:i + batch_size].contiguous()
        ax, co, sa, at = ctx.gather_patches(vol, c, atlas=d_atlas, bg_fix=True)
        yield (ax.cpu().numpy().astype(datatype, copy=False), co.cpu().numpy().astype(datatype, copy=False),
               sa.cpu().numpy().astype(datatype, copy=False), at.cpu().numpy(), [tuple(int(v) for v in r) for r in c.cpu().numpy()])


def normalise_test(image):
    """(image - mean_nz) / std_nz with numpy's own dtype promotion. [base.py:358]"""
    nz = image[np.nonzero(image)]
    return (image - nz.mean()) / nz.std()


def _crop_enabled(options):
    crop = options.get('crop_bool')
    if crop is None:
        crop = str(options.get('crop', 'True')).strip().lower() in ('true', '1', 'yes', 'on')
    return bool(crop)


def candidate_mask(image, dir_name, options):
    """bool volume of the voxels to classify. [base.py:367-372]"""
    if _crop_enabled(options):
        mask_atlas = nib.load(os.path.join(dir_name, 'tmp', 'MNI_subcortical_mask.nii.gz')).get_data()
        ctx = get_context(options.get('device'))
        m = _dev(np.ascontiguousarray(mask_atlas != 0).view(np.uint8), ctx.device)
        return ctx.dilate_mask(m, 10).cpu().numpy().astype(bool)   # == ndimage.binary_dilation(mask, iterations=10)
    return image.astype(bool)


def _require_registered(atlas_name):
    if not os.path.exists(atlas_name):
        raise IOError("%s is missing: atlas registration (register_masks / niftyreg) is outside this "
                      "implementation's scope -- create the tmp/MNI_* files with the reference pipeline" % atlas_name)


# ---------------------------------------------------------------------------------------------
# inference
# ---------------------------------------------------------------------------------------------
def test_scan(net, test_scan, options):
    """Segment one scan and write the reference's output files; returns elapsed minutes.
    [base.py:401-458]

    options['inference'] = 'dense' (default): whole-volume dilated formulation restricted to
    the bounding box of the candidates (sc_segment_volume); 'patchwise': the reference's own
    flow, batch by batch through net.predict / predict_proba on gathered patches.
    """
    s_time = time.time()
    image_path, name = os.path.split(test_scan)
    mode = options.get('inference', 'dense')
    # I/O: in the dense mode the files are read straight into page-locked memory, in the array order they have on disk
    t1_nii = nib.load(test_scan, pinned=(mode != 'patchwise'))
    t1 = t1_nii.get_data()
    want_proba = options['out_probabilities'] == 'True'
    if mode == 'patchwise':
        image = np.zeros_like(t1)
        image_proba = np.zeros(t1_nii.shape + (15,)) if want_proba else None
        for batch_axial, batch_cor, batch_sag, atlas, centers in load_patch_batch(test_scan, options):
            X = {'in1': batch_axial, 'in2': batch_cor, 'in3': batch_sag, 'in4': atlas}
            x, y, z = np.stack(centers, axis=1)
            if want_proba:
                y_pred_proba = net.predict_proba(X)
                image[x, y, z] = np.argmax(y_pred_proba, axis=1)
                image_proba[x, y, z, :] = y_pred_proba
            else:
                image[x, y, z] = net.predict(X)
    else:
        atlas_name = os.path.join(image_path, 'tmp', 'MNI_sub_probabilities.nii.gz')
        _require_registered(atlas_name)
        atlas = nib.load(atlas_name, pinned=True).get_data()
        crop_mask = None
        if _crop_enabled(options):
            crop_mask = nib.load(os.path.join(image_path, 'tmp', 'MNI_subcortical_mask.nii.gz'), pinned=True).get_data()
        timings = options.get('timings')
        post_mask = None
        if options['post_process'] == 'True':      # filtered on the device, before the label volume is downloaded
            post_mask = crop_mask if crop_mask is not None else nib.load(os.path.join(image_path, 'tmp', 'MNI_subcortical_mask.nii.gz'), pinned=True).get_data()
        labels, image_proba, n_cand = segment_arrays(net.ctx, t1, atlas, crop_mask, want_proba, timings, post_mask=post_mask)
        if options['debug'] == 'True':
            print("    -->  num of samples to test:", n_cand)
        image = labels.astype(t1.dtype)

    if want_proba:
        nib.Nifti1Image(image_proba, affine=t1_nii.affine).to_filename(os.path.join(image_path, 'out_subcortical_prob.nii.gz'))
    if options['post_process'] == 'True':
        filtered = image if mode != 'patchwise' else post_process_segmentation(image_path, image, device=options.get('device'))
        nib.Nifti1Image(filtered, affine=t1_nii.affine).to_filename(os.path.join(image_path, 'out_subcortical_seg_prec.nii.gz'))
    else:
        nib.Nifti1Image(image, affine=t1_nii.affine).to_filename(os.path.join(image_path, 'out_subcortical_rawseg.nii.gz'))
    return (time.time() - s_time) / 60.0


def segment_arrays(ctx, t1, atlas, crop_mask=None, want_proba=False, timings=None, post_mask=None):
    """The timed part of test_scan for one scan, device-resident from the first byte [base.py:357-372, 401-440]:
    raw T1 / atlas priors (/ registered mask) as host arrays in -> uint8 label volume (+ float32 [X,Y,Z,15] probabilities,
    number of candidates) out.  One upload per array (Fortran-ordered NIfTI arrays are reordered on the device), then
    normalisation over the non-zero voxels (numpy's result bit for bit), candidate mask (non-zero T1 voxels, or the
    registered mask dilated 10 times), its bounding box, the dense network pass and the scatter, all without the mask,
    the normalised volume or the coordinates ever visiting the host; one download of the result."""
    import torch
    t0 = time.time()
    shape = tuple(int(s) for s in t1.shape[:3])
    atlas = np.asarray(atlas)
    if atlas.dtype != np.float32:          # a scaled NIfTI comes back as float64: the priors are consumed as float32 (base.py:388)
        atlas = atlas.astype(np.float32, order='K')
    main = torch.cuda.current_stream()
    side = _side_stream(ctx.device)
    atlas_ready = torch.cuda.Event()
    d_atlas = None

    def start_atlas_upload(box):
        # The priors (94 % of the upload) are first needed by the FC head: they are uploaded and reordered on a side stream while
        # the convolution phase runs -- enqueued behind the host-synchronous steps that precede it, whose small device -> host
        # copies would otherwise queue up behind the gigabyte.  Crop mode reads the priors at the candidates only: just the
        # bounding box of the dilated mask is uploaded (strided DMA from the host array).
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if box is not None:
                d = ctx.upload_volume_box(atlas, box, channels=15)
            else:
                d, _ = ctx.upload_volume(atlas, channels=15)
                d = d.view(torch.float32).view(shape + (15,))
            atlas_ready.record(side)
        return d

    # T1 first: copies of one direction are served in issue order, and the convolution phase only waits for the T1 volume
    raw, dt = ctx.upload_volume(t1)
    vol, mean, std = ctx.normalise_volume(raw, dt, shape)
    if crop_mask is not None:
        mraw, mdt = ctx.upload_volume(crop_mask)
        cand = ctx.dilate_mask(ctx.candidate_mask(mraw, mdt, shape), 10)   # == ndimage.binary_dilation(mask, iterations=10)
    else:
        cand = ctx.candidate_mask(raw, dt, shape)                          # == image.astype('bool')
    box, n_cand = ctx.mask_bbox(cand)
    if box is not None:
        d_atlas = start_atlas_upload(box if crop_mask is not None else None)
    lab = torch.zeros(shape, dtype=torch.uint8, device=vol.device)
    prob = torch.zeros(shape + (15,), dtype=torch.float32, device=vol.device) if want_proba else None
    if box is not None:
        ctx.segment_volume(vol, d_atlas, box=box, cand_mask=cand, label_vol=lab, proba_vol=prob, atlas_ready=atlas_ready)
    main.wait_stream(side)
    if post_mask is not None:        # post_process_segmentation (base.py:460-480) on the device-resident label volume
        if post_mask is crop_mask and crop_mask is not None:
            pm = ctx.candidate_mask(mraw, mdt, shape)
        else:
            praw, pdt = ctx.upload_volume(post_mask)
            pm = ctx.candidate_mask(praw, pdt, shape)
        lab = ctx.post_process(lab, pm)
    h_lab = _pinned_out('lab', shape, torch.uint8)
    h_lab.copy_(lab, non_blocking=True)
    h_prob = None
    if want_proba:
        h_prob = _pinned_out('prob', shape + (15,), torch.float32)
        h_prob.copy_(prob, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    if timings is not None:
        timings.update({'hot_s': time.time() - t0, 'n_candidates': n_cand, 'mean_nz': mean, 'std_nz': std, 'box': box})
    # views of reusable page-locked buffers: valid until the next call
    return h_lab.numpy(), (h_prob.numpy() if want_proba else None), n_cand


_pinned_cache = {}
_side_streams = {}


def _side_stream(device):
    import torch
    if device not in _side_streams:
        _side_streams[device] = torch.cuda.Stream(device=device)
    return _side_streams[device]


def _pinned_out(key, shape, dtype):
    """reusable page-locked result buffers (allocating 1 GB of pinned memory per scan would cost more than the scan)"""
    import torch
    t = _pinned_cache.get(key)
    n = int(np.prod(shape))
    if t is None or t.dtype != dtype or t.numel() < n:
        t = torch.empty(n, dtype=dtype, pin_memory=True)
        _pinned_cache[key] = t
    return t[:n].view(shape)


def bounding_box(mask):
    """half-open (x0, x1, y0, y1, z0, z1) of the True voxels, or None"""
    if not mask.any():
        return None
    box = []
    for ax in range(3):
        other = tuple(a for a in range(3) if a != ax)
        nz = np.nonzero(mask.any(axis=other))[0]
        box += [int(nz[0]), int(nz[-1]) + 1]
    return tuple(box)


def post_process_segmentation(image_folder, input_mask, device=None):
    """Per label 1..14 keep the 6-connected component overlapping the registered sub-cortical mask most. [base.py:460-480]
    Runs on the device (sc_post_process: one union-find labelling pass for all classes), with the reference's quirks: a
    class without any component inside the mask selects "component 0", i.e. everything that is not of that class."""
    import torch
    ctx = get_context(device)
    atlas = load_nii(os.path.join(image_folder, 'tmp', 'MNI_subcortical_mask.nii.gz')).get_data()
    shape = tuple(int(s) for s in input_mask.shape[:3])
    seg = torch.from_numpy(np.ascontiguousarray(input_mask).astype(np.uint8)).to('cuda:%d' % ctx.device)
    mraw, mdt = ctx.upload_volume(atlas)
    mask = ctx.candidate_mask(mraw, mdt, shape)
    out = ctx.post_process(seg, mask)
    return out.cpu().numpy().astype(input_mask.dtype)


def register_masks(input_mask):
    raise NotImplementedError("atlas registration (niftyreg reg_aladin / reg_f3d / reg_resample, base.py:483-551) "
                              "stays with the reference pipeline; it is outside the hot path rebuilt here")


# ---------------------------------------------------------------------------------------------
# training data
# ---------------------------------------------------------------------------------------------
def load_data(options):
    """-> (x_axial, x_cor, x_sag, y_axial, x_atlas, names), lists indexed by subject.
    [base.py:11-37]"""
    (x_axial, y_axial, x_cor, y_cor, x_sag, y_sag, x_atlas, names) = load_patches(
        dir_name=options['train_folder'], t1_name=options['t1_name'], mask_name=options['roi_name'],
        size=tuple(options['patch_size']), device=options.get('device'))
    return x_axial, x_cor, x_sag, y_axial, x_atlas, names


def load_patches(dir_name, mask_name, t1_name, size, seeds=None, balance_neg=True, device=None):
    """[base.py:221-256]"""
    print('    --> Loading ' + t1_name + ' images')
    x_axial, y_axial, x_cor, y_cor, x_sag, y_sag, centers, t1_names = load_patch_vectors(
        t1_name, mask_name, dir_name, size, balance_neg=balance_neg, device=device)
    x_atlas = get_atlas_vectors(dir_name, centers, t1_names, device=device)
    return x_axial, y_axial, x_cor, y_cor, x_sag, y_sag, x_atlas, t1_names


def load_patch_vectors(name, label_name, dir_name, size, random_state=42, balance_neg=True, device=None):
    """Boundary-restricted sampling per subject: every voxel with label 1..14 plus as many
    shuffled label-15 voxels; T1 patches of the three views and the label patches.
    [base.py:120-184]

    The reference gathers label patches for all three views only to read their centre pixel
    later (base.py:85); the returned y_* arrays are therefore [n, 32, 32] uint8 patches whose
    centre pixel holds the label gathered on the GPU and zeros elsewhere.
    """
    if tuple(size) != (32, 32):
        raise ValueError("only patch_size (32, 32) is implemented")
    ctx = get_context(device)
    subjects = [f for f in sorted(os.listdir(dir_name)) if os.path.isdir(os.path.join(dir_name, f))]
    image_names = [os.path.join(dir_name, subject, name) for subject in subjects]
    label_names = [os.path.join(dir_name, subject, label_name) for subject in subjects]
    x_axial, y_axial, x_cor, y_cor, x_sag, y_sag, vox_positions = [], [], [], [], [], [], []
    for image_name, lab_name in zip(image_names, label_names):
        im = load_nii(image_name).get_data()
        nz = im[np.nonzero(im)]
        # base.py:146 under the reference's numpy 1.12: a float32 array combined with float64 *scalars* stays float32
        # (value-based casting), i.e. the mean and std are rounded to float32 first; NumPy 2 would promote to float64
        im_norm = (im.astype(np.float32) - np.float32(nz.mean())) / np.float32(nz.std())
        mask = load_nii(lab_name).get_data()
        pos = get_mask_voxels(np.logical_and(mask > 0, mask < 15), device=ctx.device, as_array=True)
        neg = get_mask_voxels(mask == 15, size=len(pos) if balance_neg else None, device=ctx.device, as_array=True)
        cen = np.concatenate([pos, neg]).astype(np.int32)
        vol = _dev(im_norm, ctx.device, np.float32)
        cdev = _centers_tensor(cen, ctx.device)
        ax, co, sa, _ = ctx.gather_patches(vol, cdev)
        lab = ctx.gather_center_labels(_dev(mask, ctx.device, np.uint8), cdev).cpu().numpy()
        ypatch = np.zeros((len(cen), 32, 32), np.uint8)
        ypatch[:, 16, 16] = lab
        x_axial.append(ax[:, 0].cpu().numpy()); x_cor.append(co[:, 0].cpu().numpy()); x_sag.append(sa[:, 0].cpu().numpy())
        y_axial.append(ypatch); y_cor.append(ypatch); y_sag.append(ypatch)
        vox_positions.append(cen)
    return x_axial, y_axial, x_cor, y_cor, x_sag, y_sag, vox_positions, image_names


def get_atlas_vectors(dir_name, centers, t1_names, device=None):
    """Atlas prior vector at every training centre, no background fix (the reference's fix
    there never fires, quirk Q4). [base.py:187-218]"""
    ctx = get_context(device)
    subjects = [f for f in sorted(os.listdir(dir_name)) if os.path.isdir(os.path.join(dir_name, f))]
    atlas_names = [os.path.join(dir_name, subject, 'tmp', 'MNI_sub_probabilities.nii.gz') for subject in subjects]
    out = []
    for atlas_name, c in zip(atlas_names, centers):
        _require_registered(atlas_name)
        atlas = load_nii(atlas_name).get_data()
        a = _dev(atlas, ctx.device, np.float32)
        dummy = a[..., 0].contiguous()
        res = ctx.gather_patches(dummy, _centers_tensor(c, ctx.device), atlas=a, bg_fix=False, views=(False, False, False))
        out.append(res[3].cpu().numpy().astype(atlas.dtype, copy=False))
    return out


def generate_training_set(x_axial, x_coronal, x_saggital, x_atlas, y, options, randomize=True):
    """Concatenate subjects, take the centre label, map 15 -> 0, shuffle everything with one
    seed, add the channel axis. [base.py:53-117]  Host-side: it is index bookkeeping."""
    x_train_axial = np.concatenate(x_axial, axis=0).astype('float32')
    x_train_cor = np.concatenate(x_coronal, axis=0).astype('float32')
    x_train_sag = np.concatenate(x_saggital, axis=0).astype('float32')
    x_train_atlas = np.concatenate(x_atlas, axis=0).astype('float32')
    y_train = np.concatenate(y, axis=0).astype('uint8')
    y_train = np.squeeze(y_train[:, y_train.shape[1] // 2, y_train.shape[2] // 2])
    y_train[y_train == 15] = 0
    if randomize:
        seed = np.random.randint(np.iinfo(np.int32).max)
        perm = np.random.RandomState(seed).permutation(len(y_train))
        # np.random.seed(seed); np.random.permutation(a) == a[RandomState(seed).permutation(len(a))]
        x_train_axial, x_train_cor, x_train_sag = x_train_axial[perm], x_train_cor[perm], x_train_sag[perm]
        y_train, x_train_atlas = y_train[perm], x_train_atlas[perm]
    x_train_axial = np.expand_dims(x_train_axial, axis=1)
    x_train_cor = np.expand_dims(x_train_cor, axis=1)
    x_train_sag = np.expand_dims(x_train_sag, axis=1)
    if options['debug'] == 'True':
        print("    --> X_TRAIN: ", x_train_axial.shape[0], x_train_axial.shape)
        print("    --> Y_TRAIN POS: ", y_train[y_train > 0].shape[0])
        print("    --> Y_TRAIN NEG: ", y_train[y_train == 0].shape[0])
    return x_train_axial, x_train_cor, x_train_sag, x_train_atlas, y_train
