"""Oracle: the three-branch CNN + atlas-fused FC head, restated on torch-CPU.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- PARITY UNPINNED for this file:
the reference delegates the arithmetic to Lasagne 0.2.dev1 / Theano 0.9.0 (absent).
The graph follows ``cnn_cort/nets.py:159-231`` line by line; layer semantics are
Lasagne's documented ones:

* Conv2DLayer: valid, stride 1, flip_filters=True (true convolution), no bias
  under ``batch_norm`` (nets.py:171-177).
* BatchNormLayer params ``[beta, gamma, mean, inv_std]``, eps 1e-4, alpha 0.1;
  y = (x - mean) * (gamma * inv_std) + beta.
* ``prelu(batch_norm(conv))``: BN -> PReLU (the ReLU is replaced), one alpha per
  channel / per dense unit.
* MaxPool2DLayer pool 2 stride 2; DenseLayer flattens (C, H, W), y = x @ W + b.
* head: concat 540 -> FC1 540 -> PReLU -> concat atlas -> fc_2 270 -> PReLU ->
  out_layer 15 -> softmax (nets.py:215-231); no dropout on the atlas (:222-223).
* training: categorical cross-entropy, mean; ``lasagne.updates.adam`` lr 1e-3
  (nets.py:233-237); dropout p=.5 at ``*_l1drop``, ``f1_drop``, ``f2_drop``.
"""
import pickle
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BRANCHES = ("axial", "coronal", "saggital")
BN_EPS = 1e-4
BN_ALPHA = 0.1
CONV_CH = ((1, 20), (20, 20), (20, 40), (40, 40), (40, 60))


def load_params(path):
    """nolearn ``save_params_to`` pickle: OrderedDict{layer name -> [arrays]} (py2 proto 2)."""
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def init_params(seed=0):
    """Random parameters with the shapes / inits ``build_model`` would create
    (GlorotUniform W, zero b, alpha .25, BN beta 0 gamma 1 mean 0 inv_std 1)."""
    rng = np.random.RandomState(seed)
    P = OrderedDict()

    def glorot(shape, fan_in, fan_out):
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    for b in BRANCHES:
        for i, (ci, co) in enumerate(CONV_CH, 1):
            P["%s_ch_conv%d" % (b, i)] = [glorot((co, ci, 3, 3), ci * 9, co * 9)]
            P["%s_ch_conv%d_bn" % (b, i)] = [np.zeros(co, np.float32), np.ones(co, np.float32),
                                             np.zeros(co, np.float32), np.ones(co, np.float32)]
            P["%s_ch_prelu%d" % (b, i)] = [np.full(co, 0.25, np.float32)]
        P["%s_d1" % b] = [glorot((540, 180), 540, 180), np.zeros(180, np.float32)]
        P["%s_prelu_d1" % b] = [np.full(180, 0.25, np.float32)]
    P["FC1"] = [glorot((540, 540), 540, 540), np.zeros(540, np.float32)]
    P["prelu_f1"] = [np.full(540, 0.25, np.float32)]
    P["fc_2"] = [glorot((555, 270), 555, 270), np.zeros(270, np.float32)]
    P["prelu_f2"] = [np.full(270, 0.25, np.float32)]
    P["out_layer"] = [glorot((270, 15), 270, 15), np.zeros(15, np.float32)]
    return P


# ----------------------------------------------------------------------------
# precision emulation -- used by design-validation tests
# ----------------------------------------------------------------------------
def _round_tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _trunc_tf32(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def _q(x, mode):
    if mode == "tf32":
        return _round_tf32(x)
    if mode == "tf32t":
        return _trunc_tf32(x)
    if mode == "bf16":
        return x.bfloat16().float()
    if mode == "fp16":
        return x.half().float()
    raise ValueError(mode)


def _q8(x):
    """e4m3 (4 significant bits, |x| <= 448, saturating) -- the cross-term operands of the 'fp16+e4m3' scheme"""
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def _contract(op, x, w, mode):
    """op(x, w) with both operands rounded as ``mode`` says; '<m>x3' = hi/lo split, 3 products.

    Two-MMA candidates examined for the FC head (none is shipped):
    'fp16_w2'   single fp16 activations x fp16 hi/lo weights        (xh*wh + xh*wl)
    'fp16_x2'   fp16 hi/lo activations x single fp16 weights        (xh*wh + xl*wh)
    'fp16+e4m3' fp16 hi product + both cross terms in e4m3 with a uniform 2^-11 block scale (tcgen05 kind::mxf8f6f4)
    """
    if mode is None:
        return op(x, w)
    if mode == "fp16_w2":
        xh, wh = _q(x, "fp16"), _q(w, "fp16")
        return op(xh, wh) + op(xh, _q(w - wh, "fp16"))
    if mode == "fp16_x2":
        xh, wh = _q(x, "fp16"), _q(w, "fp16")
        return op(xh, wh) + op(_q(x - xh, "fp16"), wh)
    if mode == "fp16+e4m3":
        s = 2.0 ** 11
        xh, wh = _q(x, "fp16"), _q(w, "fp16")
        return op(xh, wh) + op(_q8(x), _q8((w - wh) * s) / s) + op(_q8((x - xh) * s) / s, _q8(w))
    if mode.endswith("x3"):
        m = mode[:-2]
        xh, wh = _q(x, m), _q(w, m)
        xl, wl = _q(x - xh, m), _q(w - wh, m)
        return op(xh, wh) + op(xh, wl) + op(xl, wh)
    return op(_q(x, mode), _q(w, mode))


def _t(a, dtype):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dtype)


def _prelu(x, alpha):
    shape = [1, -1] + [1] * (x.dim() - 2)
    return torch.where(x > 0, x, alpha.view(shape) * x)


def _bn_affine(P, name, dtype):
    beta, gamma, mean, inv_std = (_t(a, dtype) for a in P[name])
    scale = gamma * inv_std
    return scale, beta - mean * scale


def branch_forward(P, b, x, dtype=torch.float32, emulate=None, taps=None):
    """One view: [N,1,32,32] -> [N,180]  (nets.py:170-180 for axial, :186-196, :202-212)."""
    emulate = emulate or {}
    x = x.to(dtype)
    for i in range(1, 6):
        w = torch.flip(_t(P["%s_ch_conv%d" % (b, i)][0], dtype), dims=[2, 3])  # flip_filters=True
        x = _contract(F.conv2d, x, w, emulate.get("c%d" % i))
        scale, shift = _bn_affine(P, "%s_ch_conv%d_bn" % (b, i), dtype)
        x = x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        x = _prelu(x, _t(P["%s_ch_prelu%d" % (b, i)][0], dtype))
        if i in (2, 4):
            x = F.max_pool2d(x, 2)
        if taps is not None:
            taps["%s_c%d" % (b, i)] = x
    x = x.flatten(1)
    W, bias = (_t(a, dtype) for a in P["%s_d1" % b])
    x = _contract(torch.matmul, x, W, emulate.get("d1")) + bias
    return _prelu(x, _t(P["%s_prelu_d1" % b][0], dtype))


def head_forward(P, feats, atlas, dtype=torch.float32, emulate=None, logits=False):
    """[N,540] features + [N,15] atlas -> softmax [N,15]  (nets.py:215-231)."""
    emulate = emulate or {}
    W, bias = (_t(a, dtype) for a in P["FC1"])
    x = _prelu(_contract(torch.matmul, feats, W, emulate.get("fc1")) + bias, _t(P["prelu_f1"][0], dtype))
    x = torch.cat([x, atlas.to(dtype)], dim=1)
    W, bias = (_t(a, dtype) for a in P["fc_2"])
    x = _prelu(_contract(torch.matmul, x, W, emulate.get("fc2")) + bias, _t(P["prelu_f2"][0], dtype))
    W, bias = (_t(a, dtype) for a in P["out_layer"])
    z = _contract(torch.matmul, x, W, emulate.get("out")) + bias
    return z if logits else torch.softmax(z, dim=1)


def forward(P, in1, in2, in3, in4, dtype=torch.float32, emulate=None, minibatch=128):
    """Deterministic ``predict_proba`` on host arrays, in nolearn-sized minibatches."""
    outs = []
    n = in1.shape[0]
    with torch.no_grad():
        for s in range(0, n, minibatch):
            xs = [torch.as_tensor(np.ascontiguousarray(a[s:s + minibatch])) for a in (in1, in2, in3)]
            feats = torch.cat([branch_forward(P, b, x, dtype, emulate) for b, x in zip(BRANCHES, xs)], dim=1)
            at = torch.as_tensor(np.ascontiguousarray(in4[s:s + minibatch]))
            outs.append(head_forward(P, feats, at, dtype, emulate))
    if not outs:
        return np.zeros((0, 15), np.float32)
    return torch.cat(outs).numpy()


def predict(P, in1, in2, in3, in4, **kw):
    """nolearn ``predict`` = argmax of ``predict_proba`` (first maximum wins)."""
    return np.argmax(forward(P, in1, in2, in3, in4, **kw), axis=1)


# ----------------------------------------------------------------------------
# dense dilated reformulation: one whole slice per view
# ----------------------------------------------------------------------------
def dense_branch(P, b, sl, dtype=torch.float64):
    """Slice [H,W] -> per-pixel branch features [180,H,W]; identical to running
    :func:`branch_forward` on the 32x32 patch around every pixel (zeros outside)."""
    x = F.pad(torch.as_tensor(sl).to(dtype)[None, None], (16, 15, 16, 15))
    dil = (1, 1, 2, 2, 4)
    for i in range(1, 6):
        w = torch.flip(_t(P["%s_ch_conv%d" % (b, i)][0], dtype), dims=[2, 3])
        x = F.conv2d(x, w, dilation=dil[i - 1])
        scale, shift = _bn_affine(P, "%s_ch_conv%d_bn" % (b, i), dtype)
        x = x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        x = _prelu(x, _t(P["%s_ch_prelu%d" % (b, i)][0], dtype))
        if i == 2:
            x = F.max_pool2d(x, 2, stride=1, dilation=1)
        if i == 4:
            x = F.max_pool2d(x, 2, stride=1, dilation=2)
    W, bias = (_t(a, dtype) for a in P["%s_d1" % b])
    w = W.t().reshape(180, 60, 3, 3)  # dense layer: flatten order (c, h, w), no flip
    x = F.conv2d(x, w, dilation=4) + bias.view(1, -1, 1, 1)
    return _prelu(x, _t(P["%s_prelu_d1" % b][0], dtype))[0]


def dense_volume_features(P, vol, dtype=torch.float64):
    """[X,Y,Z] -> [X,Y,Z,540] branch features (axial | coronal | saggital)."""
    X, Y, Z = vol.shape
    out = torch.zeros((X, Y, Z, 540), dtype=dtype)
    with torch.no_grad():
        for z in range(Z):  # axial: patch axes (x, y) at fixed z
            out[:, :, z, 0:180] = dense_branch(P, "axial", vol[:, :, z], dtype).permute(1, 2, 0)
        for y in range(Y):  # coronal: (x, z) at fixed y
            out[:, y, :, 180:360] = dense_branch(P, "coronal", vol[:, y, :], dtype).permute(1, 2, 0)
        for x in range(X):  # saggital: (y, z) at fixed x
            out[x, :, :, 360:540] = dense_branch(P, "saggital", vol[x, :, :], dtype).permute(1, 2, 0)
    return out


def dense_volume_forward(P, vol, atlas_fixed, dtype=torch.float64):
    """Whole small volume -> softmax [X,Y,Z,15]; ``atlas_fixed`` already has the bg-fix."""
    feats = dense_volume_features(P, vol, dtype).reshape(-1, 540)
    with torch.no_grad():
        p = head_forward(P, feats, torch.as_tensor(atlas_fixed).reshape(-1, 15), dtype)
    return p.reshape(vol.shape + (15,)).numpy()


# ----------------------------------------------------------------------------
# training step (nolearn train_fn): BN batch statistics, dropout, CE, Lasagne Adam
# ----------------------------------------------------------------------------
def trainable_names(P):
    """(layer name, array index) of every trainable array in pickle order: conv W,
    BN beta/gamma, PReLU alpha, dense W/b.  BN mean / inv_std (index 2, 3) are state."""
    out = []
    for name, arrs in P.items():
        for k in range(len(arrs)):
            if name.endswith("_bn") and k >= 2:
                continue
            out.append((name, k))
    return out


def train_forward(T, in1, in2, in3, in4, y, masks=None, dtype=torch.float64):
    """Training-mode forward on torch parameters ``T`` (dict name -> list of tensors).

    Returns (mean CE loss, dict of BN batch (mean, inv_std), softmax).  ``masks``:
    dict {'axial_l1drop','coronal_l1drop','saggital_l1drop' [N,60,3,3]; 'f1_drop' [N,540];
    'f2_drop' [N,540]} of 0/1 keep masks; kept units are scaled by 1/(1-p) = 2.
    """
    stats = {}
    feats = []
    for b, x in zip(BRANCHES, (in1, in2, in3)):
        x = x.to(dtype)
        for i in range(1, 6):
            w = torch.flip(T["%s_ch_conv%d" % (b, i)][0], dims=[2, 3])
            x = F.conv2d(x, w)
            beta, gamma = T["%s_ch_conv%d_bn" % (b, i)][:2]
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            inv_std = 1.0 / torch.sqrt(var + BN_EPS)
            stats["%s_ch_conv%d_bn" % (b, i)] = (mean.detach(), inv_std.detach())
            x = (x - mean.view(1, -1, 1, 1)) * (gamma * inv_std).view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
            x = _prelu(x, T["%s_ch_prelu%d" % (b, i)][0])
            if i in (2, 4):
                x = F.max_pool2d(x, 2)
        if masks is not None:
            x = x * masks["%s_l1drop" % b].to(dtype) * 2.0
        x = x.flatten(1) @ T["%s_d1" % b][0] + T["%s_d1" % b][1]
        feats.append(_prelu(x, T["%s_prelu_d1" % b][0]))
    x = torch.cat(feats, dim=1)
    if masks is not None:
        x = x * masks["f1_drop"].to(dtype) * 2.0
    x = _prelu(x @ T["FC1"][0] + T["FC1"][1], T["prelu_f1"][0])
    if masks is not None:
        x = x * masks["f2_drop"].to(dtype) * 2.0
    x = torch.cat([x, in4.to(dtype)], dim=1)
    x = _prelu(x @ T["fc_2"][0] + T["fc_2"][1], T["prelu_f2"][0])
    z = x @ T["out_layer"][0] + T["out_layer"][1]
    logp = torch.log_softmax(z, dim=1)
    loss = -logp[torch.arange(z.shape[0]), y.long()].mean()
    return loss, stats, torch.exp(logp)


def to_torch(P, dtype=torch.float64, requires_grad=True):
    T = OrderedDict()
    train = set(trainable_names(P))
    for name, arrs in P.items():
        T[name] = [_t(a, dtype).requires_grad_(requires_grad and (name, k) in train) for k, a in enumerate(arrs)]
    return T


def adam_update(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """``lasagne.updates.adam``: a_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= a_t*m/(sqrt(v)+eps)."""
    a_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    return p - a_t * m / (torch.sqrt(v) + eps), m, v


def train_step(P, in1, in2, in3, in4, y, masks=None, state=None, lr=1e-3, dtype=torch.float64):
    """One minibatch update.  Returns (loss, grads{(name,k): ndarray}, new P, new state).

    ``state`` = {'t': int, 'm': {...}, 'v': {...}} (zeros at t=0).  BN running statistics
    move by ``s <- 0.9 s + 0.1 batch`` on mean and on inv_std (Lasagne BatchNormLayer).
    """
    T = to_torch(P, dtype)
    names = trainable_names(P)
    args = [torch.as_tensor(np.ascontiguousarray(a)) for a in (in1, in2, in3, in4)]
    yt = torch.as_tensor(np.ascontiguousarray(y).astype(np.int64))
    mk = None if masks is None else {k: torch.as_tensor(np.ascontiguousarray(v)) for k, v in masks.items()}
    loss, stats, _ = train_forward(T, *args, yt, masks=mk, dtype=dtype)
    grads = torch.autograd.grad(loss, [T[n][k] for n, k in names])
    if state is None:
        state = {"t": 0, "m": {}, "v": {}}
    t = state["t"] + 1
    newP = OrderedDict((n, [np.array(a, copy=True) for a in arrs]) for n, arrs in P.items())
    new_state = {"t": t, "m": {}, "v": {}}
    G = {}
    for (n, k), g in zip(names, grads):
        p = T[n][k].detach()
        m = state["m"].get((n, k), torch.zeros_like(p))
        v = state["v"].get((n, k), torch.zeros_like(p))
        p2, m2, v2 = adam_update(p, g, m, v, t, lr=lr)
        newP[n][k] = p2.numpy().astype(np.float32)
        new_state["m"][(n, k)], new_state["v"][(n, k)] = m2, v2
        G[(n, k)] = g.numpy()
    for n, (mean, inv_std) in stats.items():
        newP[n][2] = ((1 - BN_ALPHA) * _t(P[n][2], dtype) + BN_ALPHA * mean).numpy().astype(np.float32)
        newP[n][3] = ((1 - BN_ALPHA) * _t(P[n][3], dtype) + BN_ALPHA * inv_std).numpy().astype(np.float32)
    return float(loss.detach()), G, newP, new_state
