"""Minimal NIfTI-1 single-file (.nii / .nii.gz) reader / writer.

The reference does all image I/O through nibabel 2.1.0 (``cnn_cort/base.py:145,357,413,
446-455``), which is not installable here.  This module implements the small subset the
hot-path callers need -- ``load(path).get_data()``, ``.affine``, ``.shape`` and
``Nifti1Image(data, affine).to_filename(path)`` -- and stays on the host, outside the timed
path (north_star: "NIfTI I/O ... stay on the host").
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
           256: np.int8, 512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


def _open(path, mode):
    if str(path).endswith(".gz"):
        # nibabel writes .nii.gz at compresslevel 1 (its Opener default); 9 would cost minutes on a 1 GB prior volume
        return gzip.open(path, mode, compresslevel=1) if "w" in mode else gzip.open(path, mode)
    return open(path, mode)


def _quat_affine(b, c, d, qx, qy, qz, dx, dy, dz, qfac):
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    A = np.eye(4)
    A[:3, :3] = R * np.array([dx, dy, dz * (-1.0 if qfac < 0 else 1.0)])
    A[:3, 3] = (qx, qy, qz)
    return A


class Nifti1Image(object):
    def __init__(self, data, affine=None, header=None):
        self._data = np.asarray(data)
        self.affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        self.header = header or {}

    @property
    def shape(self):
        return self._data.shape

    def get_data(self):
        return self._data

    get_fdata = get_data

    def to_filename(self, path):
        data = self._data
        if data.dtype == np.bool_:
            data = data.astype(np.uint8)
        code = _CODES.get(data.dtype.str[1:])
        if code is None:
            data = data.astype(np.float32)
            code = 16
        hdr = bytearray(348)
        struct.pack_into("<i", hdr, 0, 348)
        dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
        struct.pack_into("<8h", hdr, 40, *dim)
        struct.pack_into("<hh", hdr, 70, code, data.dtype.itemsize * 8)
        zooms = np.sqrt((self.affine[:3, :3] ** 2).sum(0))
        pixdim = [1.0] + [float(z) if z > 0 else 1.0 for z in zooms] + [1.0] * 4
        struct.pack_into("<8f", hdr, 76, *pixdim)
        struct.pack_into("<f", hdr, 108, 352.0)
        struct.pack_into("<ff", hdr, 112, 1.0, 0.0)
        hdr[123] = 10  # xyzt_units: mm + sec
        struct.pack_into("<hh", hdr, 252, 0, 2)  # qform unknown, sform aligned
        for r in range(3):
            struct.pack_into("<4f", hdr, 280 + 16 * r, *[float(v) for v in self.affine[r]])
        hdr[344:348] = b"n+1\0"
        with _open(path, "wb") as f:
            f.write(bytes(hdr))
            f.write(b"\0\0\0\0")
            f.write(np.asfortranarray(data).tobytes(order="F"))


_PINNED_KEEPALIVE = {}


def _pinned_bytes(n):
    """page-locked host buffer of n bytes as a numpy uint8 array (None when CUDA is not usable)"""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        t = torch.empty(int(n), dtype=torch.uint8, pin_memory=True)
        a = t.numpy()
        _PINNED_KEEPALIVE[a.__array_interface__["data"][0]] = t      # the tensor owns the allocation
        if len(_PINNED_KEEPALIVE) > 8:
            _PINNED_KEEPALIVE.pop(next(iter(_PINNED_KEEPALIVE)))
        return a
    except Exception:
        return None


def _endianness(head, path):
    if struct.unpack_from("<i", head, 0)[0] == 348:
        return "<"
    if struct.unpack_from(">i", head, 0)[0] == 348:
        return ">"
    raise ValueError("%s: not a NIfTI-1 file" % path)


def load(path, pinned=False):
    """pinned=True reads the file straight into page-locked memory (when CUDA is available), so that the later
    host-to-device copy of the scan runs at full PCIe speed without a staging copy; the array is otherwise identical."""
    with _open(path, "rb") as f:
        head = f.read(352)
        if len(head) < 348:
            raise ValueError("%s: not a NIfTI-1 file" % path)
        end = _endianness(head, path)
        raw = None
        if pinned:
            dim = struct.unpack_from(end + "8h", head, 40)
            code = struct.unpack_from(end + "h", head, 70)[0]
            off = int(struct.unpack_from(end + "f", head, 108)[0])
            if code in _DTYPES and 1 <= dim[0] <= 7 and off >= len(head):
                nbytes = int(np.prod([int(d) for d in dim[1:1 + dim[0]]])) * np.dtype(_DTYPES[code]).itemsize
                buf = _pinned_bytes(off + nbytes)
                if buf is not None:
                    buf[:len(head)] = np.frombuffer(head, np.uint8)
                    view = memoryview(buf)[len(head):]
                    got = 0
                    while got < len(view):
                        k = f.readinto(view[got:])
                        if not k:
                            break
                        got += k
                    raw = buf
        if raw is None:
            raw = head + f.read()
    if bytes(raw[344:347]) not in (b"n+1",):
        raise ValueError("%s: only single-file NIfTI-1 (magic n+1) is supported" % path)
    dim = struct.unpack_from(end + "8h", raw, 40)
    ndim = dim[0]
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    while len(shape) > 3 and shape[-1] == 1:
        shape = shape[:-1]
    code, _bitpix = struct.unpack_from(end + "hh", raw, 70)
    if code not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, code))
    pixdim = struct.unpack_from(end + "8f", raw, 76)
    vox_offset = int(struct.unpack_from(end + "f", raw, 108)[0])
    slope, inter = struct.unpack_from(end + "ff", raw, 112)
    qcode, scode = struct.unpack_from(end + "hh", raw, 252)
    dt = np.dtype(_DTYPES[code]).newbyteorder(end)
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=vox_offset).reshape(shape, order="F")
    if end == ">":
        data = data.astype(dt.newbyteorder("<"))
    # nibabel: scl_slope == 0 (or non-finite) means "no scaling" and scl_inter is then ignored; a non-finite inter counts as 0
    if np.isfinite(slope) and slope != 0.0:
        inter = inter if np.isfinite(inter) else 0.0
        if slope != 1.0 or inter != 0.0:
            data = data * np.float64(slope) + np.float64(inter)
    if not data.flags.writeable:
        data = data.copy(order="F")      # np.frombuffer views are read-only; callers edit volumes in place
    if scode > 0:
        A = np.eye(4)
        for r in range(3):
            A[r] = struct.unpack_from(end + "4f", raw, 280 + 16 * r)
    elif qcode > 0:
        b, c, d, qx, qy, qz = struct.unpack_from(end + "6f", raw, 256)
        A = _quat_affine(b, c, d, qx, qy, qz, pixdim[1], pixdim[2], pixdim[3], pixdim[0])
    else:
        A = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    return Nifti1Image(data, A, {"pixdim": pixdim, "datatype": code})


load_nii = load
