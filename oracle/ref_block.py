"""Oracle: conv-BN-LeakyReLU cells and YOLODetectionBlockV3 (numpy, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned by the reference (MXNet cannot run here).

Follows:
  * models/definitions/layers.py:63-70   _conv2d: Conv2D(channel, k, stride, pad, no bias) + BatchNorm(eps 1e-5) + LeakyReLU(0.1)
  * models/definitions/layers.py:73-79   _conv3d: the same with Conv3D
  * models/definitions/layers.py:82-89   _conv21d: _conv3d(m,(1,d,d)) then _conv3d(channel,(t,1,1)); Conv('21') passes m=channel (:154-156)
  * models/definitions/yolo/yolo3_temporal.py:198-239  YOLODetectionBlockV3: body / tip wiring, swapaxes(1,2) around the 3-D convs
MXNet convolution = cross-correlation: y[o,t,y,x] = sum_{c,i,j,k} w[o,c,i,j,k] * x[c, t+i-pt, y+j-ph, x+k-pw], zeros outside.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def conv_bn_lrelu(x, weight, gamma, beta, mean, var, eps=1e-5, slope=0.1):
    """x (B,T,C,H,W) [or (B,C,H,W)], weight (Co,C,kt,kh,kw) [or (Co,C,kh,kw)], 'same' zero padding, stride 1."""
    x = np.asarray(x, f32)
    w = np.asarray(weight, f32)
    four = x.ndim == 4
    if four:
        x = x[:, None]
    if w.ndim == 4:
        w = w[:, :, None]
    B, T, C, H, W = x.shape
    Co, Ci, kt, kh, kw = w.shape
    assert Ci == C
    pt, ph, pw = kt // 2, kh // 2, kw // 2
    xp = np.zeros((B, T + 2 * pt, C, H + 2 * ph, W + 2 * pw), f32)
    xp[:, pt:pt + T, :, ph:ph + H, pw:pw + W] = x
    y = np.zeros((B, T, Co, H, W), f32)
    for i in range(kt):
        for j in range(kh):
            for k in range(kw):
                patch = xp[:, i:i + T, :, j:j + H, k:k + W]                     # (B,T,C,H,W)
                y += np.einsum("oc,btchw->btohw", w[:, :, i, j, k], patch, optimize=True).astype(f32)
    scale = (np.asarray(gamma, f32) / np.sqrt(np.asarray(var, f32) + f32(eps))).astype(f32)
    sh = (1, 1, Co, 1, 1)
    y = (y - np.asarray(mean, f32).reshape(sh)) * scale.reshape(sh) + np.asarray(beta, f32).reshape(sh)
    y = np.where(y > 0, y, y * f32(slope)).astype(f32)
    return y[:, 0] if four else y


def detection_block(x, cells, conv_type="2", round_fn=None):
    """YOLODetectionBlockV3.hybrid_forward (yolo3_temporal.py:229-239).  `cells` = list of dicts
    {weight,gamma,beta,mean,var} in execution order: [1x1, expand] x 2, 1x1, then the tip's expand; an expand of
    conv_type '21' is two consecutive entries ((1,3,3) then (3,1,1)).  round_fn (e.g. bf16 rounding) is applied to
    every cell output to model the carrier precision of the device path.  Returns (route, tip)."""
    rf = round_fn if round_fn is not None else (lambda a: a)
    it = iter(cells)

    def run(z, n):
        for _ in range(n):
            c = next(it)
            z = rf(conv_bn_lrelu(z, c["weight"], c["gamma"], c["beta"], c["mean"], c["var"]))
        return z

    n_exp = 2 if conv_type == "21" else 1
    z = np.asarray(x, f32)
    for _ in range(2):
        z = run(z, 1)
        z = run(z, n_exp)
    route = run(z, 1)
    tip = run(route, n_exp)
    return route, tip


def upsample_concat(x, route):
    """yolo3.py:515-519 / yolo3_temporal.py:502-506: `_upsample(x, 2)` (layers.py:10-20: repeat along W then H = nearest), `slice_like`
    crop to the route map, concat in FRONT of the route along channels (axis -3: works for (B,C,H,W) and (B,T,C,H,W))."""
    x = np.asarray(x, f32)
    up = x.repeat(2, axis=-1).repeat(2, axis=-2)
    H, W = route.shape[-2:]
    return np.concatenate([up[..., :H, :W], np.asarray(route, f32)], axis=-3)


def yolo3_neck_tips(routes, blocks, transitions, round_fn=None, conv_type="2"):
    """YOLOV3.hybrid_forward after the stages (yolo3.py:496-521) and its temporal twin (yolo3_temporal.py:448-506, t_out): routes =
    stage outputs shallow -> deep, (B,C,H,W) or (B,T,C,H,W); blocks[i] = cell dicts of the i-th (deep -> shallow)
    YOLODetectionBlockV3 (6 for conv_type '2'/'3', 9 for '21'), transitions[i] = cell dict of `_conv2d(channel,1,0,1)` (applied per
    frame through TimeDistributed in the temporal net, :495-497).  Returns the tips deep -> shallow."""
    rf = round_fn if round_fn is not None else (lambda a: a)
    rts = list(routes)[::-1]
    x, tips = np.asarray(rts[0], f32), []
    for i, cells in enumerate(blocks):
        x, tip = detection_block(x, cells, conv_type, round_fn=round_fn)
        tips.append(tip)
        if i >= len(rts) - 1:
            break
        t = transitions[i]
        lead = x.shape[:-3]
        x4 = x.reshape((-1,) + x.shape[-3:])                                 # TimeDistributed 'reshape1'
        x4 = rf(conv_bn_lrelu(x4, t["weight"], t["gamma"], t["beta"], t["mean"], t["var"]))
        x = upsample_concat(x4.reshape(lead + x4.shape[1:]), rts[i + 1])
    return tips
