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
