"""Oracle network: structural anchors and the regression pin."""
import os

import numpy as np
import pytest
import torch

from oracle import gather, network


def test_weight_file_layout(oracle_params):
    P = oracle_params
    assert len(P) == 88
    assert sum(a.size for arrs in P.values() for a in arrs) == 883455
    assert sum(P[n][k].size for n, k in network.trainable_names(P)) == 882375
    assert P["fc_2"][0].shape == (555, 270) and P["axial_d1"][0].shape == (540, 180)
    assert (P["axial_ch_conv1_bn"][3] > 0).all()  # array 4 is inv_std


def test_regression_pin(oracle_params, golden_dir):
    G = np.load(os.path.join(golden_dir, "forward_golden.npz"))
    x = [gather.get_patches(G["vol"], G["centers"], (32, 32), m).astype(np.float32)[:, None] for m in gather.VIEWS]
    p64 = network.forward(oracle_params, *x, G["atlas"], dtype=torch.float64)
    assert np.abs(p64 - G["proba64"]).max() < 1e-9
    p32 = network.forward(oracle_params, *x, G["atlas"], dtype=torch.float32)
    assert np.abs(p32 - G["proba64"]).max() < 2e-5
    assert np.allclose(p64.sum(1), 1)


def test_one_hot_atlas_drives_class(oracle_params):
    rng = np.random.RandomState(3)
    x = [rng.randn(15, 1, 32, 32).astype(np.float32) * 0.3 for _ in range(3)]
    lab = network.predict(oracle_params, *x, np.eye(15, dtype=np.float32))
    assert list(lab) == list(range(1, 15)) + [0]


def test_dense_equals_patchwise(oracle_params):
    rng = np.random.RandomState(5)
    sl = rng.randn(21, 19)
    for b, mode in zip(network.BRANCHES, gather.VIEWS):
        dense = network.dense_branch(oracle_params, b, sl, torch.float64)
        vol = {"axial": sl[:, :, None], "coronal": sl[:, None, :], "saggital": sl[None, :, :]}[mode]
        cen = gather.get_mask_voxels(np.ones(vol.shape, bool))
        pt = gather.get_patches(vol, cen, (32, 32), mode)[:, None]
        with torch.no_grad():
            ref = network.branch_forward(oracle_params, b, torch.as_tensor(pt), torch.float64)
        got = dense.permute(1, 2, 0).reshape(-1, 180)
        assert (got - ref).abs().max() < 1e-9


def test_dense_volume_forward_small(oracle_params):
    rng = np.random.RandomState(6)
    vol = rng.randn(5, 6, 4)
    atlas = rng.rand(5, 6, 4, 15).astype(np.float32)
    atlas[0, 0, 0] = 0
    cen = gather.get_mask_voxels(np.ones(vol.shape, bool))
    av = gather.atlas_vectors_test(atlas, cen)
    dense = network.dense_volume_forward(oracle_params, vol, av.reshape(5, 6, 4, 15))
    x = [gather.get_patches(vol, cen, (32, 32), m)[:, None] for m in gather.VIEWS]
    ref = network.forward(oracle_params, *x, av, dtype=torch.float64)
    assert np.abs(dense.reshape(-1, 15) - ref).max() < 1e-9


@pytest.mark.parametrize("emulate,ok", [
    ({"d1": "tf32", "fc1": "tf32", "fc2": "tf32", "out": "tf32"}, True),
    ({k: "bf16" for k in ("c1", "c2", "c3", "c4", "c5", "d1", "fc1", "fc2", "out")}, False),
])
def test_precision_contract(oracle_params, golden_dir, emulate, ok):
    """Design check: FC sites tolerate single TF32, all-bf16 does not."""
    G = np.load(os.path.join(golden_dir, "forward_golden.npz"))
    x = [gather.get_patches(G["vol"], G["centers"], (32, 32), m).astype(np.float32)[:, None] for m in gather.VIEWS]
    rng = np.random.RandomState(9)
    at = rng.dirichlet(np.ones(15) * 0.3, size=len(G["centers"])).astype(np.float32)
    ref = network.forward(oracle_params, *x, at, dtype=torch.float64)
    got = network.forward(oracle_params, *x, at, dtype=torch.float32, emulate=emulate)
    err = np.abs(got - ref).max()
    assert (err < 1e-3) == ok, err


def test_train_step_decreases_loss_and_updates_state():
    P = network.init_params(0)
    rng = np.random.RandomState(0)
    n = 12
    x = [rng.randn(n, 1, 32, 32).astype(np.float32) for _ in range(3)]
    at = rng.rand(n, 15).astype(np.float32)
    y = rng.randint(0, 15, n).astype(np.uint8)
    state, losses = None, []
    for _ in range(6):
        loss, G, P, state = network.train_step(P, *x, at, y, masks=None, state=state, lr=1e-3)
        losses.append(loss)
    assert losses[-1] < losses[0] and state["t"] == 6
    assert len(G) == len(network.trainable_names(P))
    assert not np.allclose(P["axial_ch_conv1_bn"][2], 0)  # running mean moved


def test_oracle_agrees_with_an_independent_numpy_scipy_restatement(oracle_params):
    """Second, independent restatement of nets.py:159-231 in float64 numpy: the convolutions are
    scipy.signal.convolve2d(mode='valid') -- a TRUE convolution, which is what Lasagne's flip_filters=True computes,
    so no kernel flip appears anywhere here (the torch oracle cross-correlates with flipped taps); pooling, BN, PReLU,
    the (C, H, W) flatten order, the concat order and the softmax are written out with plain numpy.  Run on the
    committed weights."""
    from scipy.signal import convolve2d
    P = oracle_params
    rng = np.random.RandomState(3)
    n = 3
    x = [rng.randn(n, 1, 32, 32).astype(np.float32).astype(np.float64) for _ in range(3)]
    atlas = rng.dirichlet(np.ones(15) * 0.4, size=n).astype(np.float32).astype(np.float64)

    def prelu(v, a):
        return np.where(v > 0, v, a * v)

    feats = []
    for b, xin in zip(network.BRANCHES, x):
        rows = []
        for s in range(n):
            maps = [xin[s, 0]]
            for i in range(1, 6):
                W = P["%s_ch_conv%d" % (b, i)][0].astype(np.float64)
                beta, gamma, mean, inv_std = [a.astype(np.float64) for a in P["%s_ch_conv%d_bn" % (b, i)]]
                alpha = P["%s_ch_prelu%d" % (b, i)][0].astype(np.float64)
                out = []
                for co in range(W.shape[0]):
                    acc = sum(convolve2d(maps[ci], W[co, ci], mode="valid") for ci in range(W.shape[1]))
                    acc = (acc - mean[co]) * (gamma[co] * inv_std[co]) + beta[co]
                    out.append(prelu(acc, alpha[co]))
                if i in (2, 4):   # MaxPool2DLayer(pool_size=2)
                    out = [o.reshape(o.shape[0] // 2, 2, o.shape[1] // 2, 2).max(axis=(1, 3)) for o in out]
                maps = out
            flat = np.stack(maps).reshape(-1)                      # (C, H, W) order
            Wd, bd = [a.astype(np.float64) for a in P["%s_d1" % b]]
            rows.append(prelu(flat @ Wd + bd, P["%s_prelu_d1" % b][0].astype(np.float64)))
        feats.append(np.stack(rows))
    h = np.concatenate(feats, 1)
    h = prelu(h @ P["FC1"][0].astype(np.float64) + P["FC1"][1], P["prelu_f1"][0].astype(np.float64))
    h = np.concatenate([h, atlas], 1)
    h = prelu(h @ P["fc_2"][0].astype(np.float64) + P["fc_2"][1], P["prelu_f2"][0].astype(np.float64))
    z = h @ P["out_layer"][0].astype(np.float64) + P["out_layer"][1]
    e = np.exp(z - z.max(1, keepdims=True))
    ref = e / e.sum(1, keepdims=True)
    x32, a32 = [a.astype(np.float32) for a in x], atlas.astype(np.float32)   # the restatement above sees the same rounded inputs
    got = network.forward(P, *x32, a32, dtype=torch.float64)
    assert np.abs(got - ref).max() < 1e-9
