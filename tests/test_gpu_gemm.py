"""K3 pieces: each dense layer of the head through both GEMM back-ends (SIMT fp32, tcgen05 bf16x3)
vs a float64 numpy product on the committed weights."""
import numpy as np
import pytest
import torch

from oracle import network as on
from gpu_util import cuda_ctx, dev

pytestmark = pytest.mark.gpu


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


def _ref(P, which, x):
    x = x.astype(np.float64)
    if which < 3:
        b = on.BRANCHES[which]
        W, bias, al = P["%s_d1" % b][0], P["%s_d1" % b][1], P["%s_prelu_d1" % b][0]
        z = x[:, :540] @ W.astype(np.float64) + bias
    elif which == 3:
        W, bias, al = P["FC1"][0], P["FC1"][1], P["prelu_f1"][0]
        z = x[:, :540] @ W.astype(np.float64) + bias
    else:
        W, bias, al = P["fc_2"][0], P["fc_2"][1], P["prelu_f2"][0]
        z = x[:, :555] @ W.astype(np.float64) + bias
    return np.where(z > 0, z, al * z)


@pytest.mark.parametrize("which", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n", [1, 128, 333, 4097])
def test_dense_layer_backends(ctx, P, which, n):
    rng = np.random.RandomState(which * 100 + n)
    width_in = 576
    k = 555 if which == 4 else 540
    x = np.zeros((n, width_in), np.float32)
    x[:, :k] = rng.randn(n, k).astype(np.float32)
    ref = _ref(P, which, x)
    if which == 3:   # FC1 reads the feature buffer: 192 columns per view, 180 used
        f = x[:, :540].copy()
        x[:] = 0
        for v in range(3):
            x[:, v * 192:v * 192 + 180] = f[:, v * 180:(v + 1) * 180]
    ncol = ref.shape[1]
    scale = np.abs(ref).max()
    got0 = ctx.dense_layer(which, dev(x), 0).cpu().numpy()[:, :ncol]
    assert np.abs(got0 - ref).max() < 2e-5 * max(1.0, scale)
    if ctx.counter("gemm") == 1:
        got1 = ctx.dense_layer(which, dev(x), 1).cpu().numpy()[:, :ncol]
        # split bf16 (hi + lo, three MMAs): ~2^-16 relative per product, K <= 555 random-sign terms
        assert np.abs(got1 - ref).max() < 1e-4 * max(1.0, scale), np.abs(got1 - ref).max()
