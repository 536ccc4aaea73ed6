import numpy as np
import pytest


def cuda_ctx():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from cnn_cort import _native
    torch.cuda.set_device(0)
    return _native.Context(0)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()
