"""Oracle: temporal pieces of the head (numpy, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned by the reference.

Follows:
  * models/definitions/layers.py:73-89   _conv3d / _conv21d: Conv3D((t,1,1), pad (1,0,0), no bias)
                                          + BatchNorm(eps 1e-5) + LeakyReLU(0.1)   (temporal cell)
  * models/definitions/layers.py:208-264  TimeDistributed, style 'reshape1'
  * models/definitions/layers.py:161-205  TemporalPooling, style 'direct'
  * models/definitions/yolo/yolo3.py:1134-1138  late 'cat' / 'max' / 'mean' joins
  * models/definitions/yolo/yolo3_temporal.py:226-239,448-468,542-555  wiring (swapaxes, t_out NMS)
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def temporal_conv_bn_lrelu(x, weight, gamma, beta, mean, var, eps=1e-5, slope=0.1):
    """Temporal (3,1,1) cell of _conv21d (layers.py:87 -> :73-79).

    x (B,T,C,H,W) [the block swaps to (B,C,T,H,W) and back, yolo3_temporal.py:231-239];
    weight (Cout,Cin,3,1,1) or (Cout,Cin,3); zero padding of 1 frame on both ends of T.
    Inference BatchNorm: (y-mean)/sqrt(var+eps)*gamma+beta, then LeakyReLU(slope).
    """
    x = np.asarray(x, f32)
    B, T, C, H, W = x.shape
    w = np.asarray(weight, f32).reshape(weight.shape[0], C, 3)
    Co = w.shape[0]
    xf = x.reshape(B, T, C, H * W)
    y = np.zeros((B, T, Co, H * W), f32)
    for t in range(T):
        for k in range(3):
            ts = t + k - 1
            if ts < 0 or ts >= T:
                continue                                  # zero padding (layers.py:87 padding=(1,0,0))
            y[:, t] += np.matmul(w[None, :, :, k], xf[:, ts])
    scale = (np.asarray(gamma, f32) / np.sqrt(np.asarray(var, f32) + f32(eps))).astype(f32)
    y = (y - np.asarray(mean, f32).reshape(1, 1, Co, 1)) * scale.reshape(1, 1, Co, 1) \
        + np.asarray(beta, f32).reshape(1, 1, Co, 1)
    y = np.where(y > 0, y, y * f32(slope)).astype(f32)
    return y.reshape(B, T, Co, H, W)


def time_distributed(fn, x):
    """TimeDistributed 'reshape1' (layers.py:241-250): (B,T,...)->(B*T,...)->fn->(B,T,...)."""
    B, T = x.shape[:2]
    y = fn(x.reshape((B * T,) + x.shape[2:]))
    if isinstance(y, (tuple, list)):
        return type(y)(yi.reshape((B, T) + yi.shape[1:]) for yi in y)
    return y.reshape((B, T) + y.shape[1:])


def temporal_pooling(x, type="max"):
    """TemporalPooling 'direct' (layers.py:202-205): max / mean over axis 1."""
    x = np.asarray(x, f32)
    return x.max(axis=1) if type == "max" else x.mean(axis=1, dtype=f32).astype(f32)


def late_cat(x):
    """k_join_type='cat', late (yolo3.py:1134-1136): (B,K,C,H,W) -> (B,K*C,H,W)."""
    B, K, C, H, W = x.shape
    return np.asarray(x, f32).reshape(B, K * C, H, W)
