"""CPU baseline of the head -> decode -> NMS path = the oracle restatement timed on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is what `bench.py`'s cpu_baseline leg
and `bench.py --impl reference` execute: MXNet cannot be installed here, so the "reference CPU head"
is the literal restatement of its path (kind = "port"):
   1x1 conv          torch.nn.functional.conv2d fp32 on CPU (oneDNN, all threads)  ~ MXNet's MKL-DNN conv  (yolo3.py:157)
   decode + concat   numpy fp32, frames split over threads, materialising (B, rows, 6) like MXNet   (yolo3.py:158-199,523)
   box_nms + slice   C restatement, images split over threads                       (yolo3.py:526-534)
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import ref_head, ref_nms


def _decode_threaded(pred, anchors, stride, num_class, threads):
    """The numpy decode split over frames on a thread pool (numpy releases the GIL inside its kernels): MXNet's CPU elementwise
    operators are OpenMP-parallel, so a single-threaded decode would understate the reference."""
    B = pred.shape[0]
    n = max(1, min(int(threads), B))
    if n == 1:
        return ref_head.decode(pred, anchors, stride, num_class)
    from concurrent.futures import ThreadPoolExecutor
    bounds = [(B * i) // n for i in range(n + 1)]
    with ThreadPoolExecutor(max_workers=n) as ex:
        parts = list(ex.map(lambda i: ref_head.decode(pred[bounds[i]:bounds[i + 1]], anchors, stride, num_class), range(n)))
    return np.concatenate(parts, axis=0)


def head_forward_cpu(tips, ws, bs, num_class, nms_thresh=0.45, valid_thresh=0.01, nms_topk=400, post_nms=100,
                     threads=None, use_torch_conv=True):
    threads = threads or os.cpu_count() or 1
    dets = []
    for t, w, b, a, s in zip(tips, ws, bs, ref_head.ANCHORS_OUT_ORDER, ref_head.STRIDES_OUT_ORDER):
        if use_torch_conv:
            import torch
            torch.set_num_threads(threads)
            with torch.no_grad():
                pred = torch.nn.functional.conv2d(torch.from_numpy(t), torch.from_numpy(w), torch.from_numpy(b)).numpy()
        else:
            pred = ref_head.conv1x1(t, w, b)
        dets.append(_decode_threaded(pred, a, s, num_class, threads))
    det = np.concatenate(dets, axis=1)
    out = ref_nms.box_nms(det, overlap_thresh=nms_thresh, valid_thresh=valid_thresh, topk=nms_topk, id_index=0,
                          score_index=1, coord_start=2, force_suppress=False, threads=threads)
    out = out[:, :post_nms]
    return out[..., 0:1], out[..., 1:2], out[..., 2:]


def time_head_cpu(tips, ws, bs, num_class, repeats=1, threads=None):
    """Returns (frames_per_second, seconds, frames) over `repeats` passes of the given sample."""
    frames = tips[0].shape[0]
    t0 = time.perf_counter()
    for _ in range(repeats):
        head_forward_cpu(tips, ws, bs, num_class, threads=threads)
    dt = time.perf_counter() - t0
    return frames * repeats / dt, dt, frames * repeats


# ------------------------------------------------------------------------------------------------
# temporal head (BASELINE configs[3]): Conv3D((3,1,1)) + BN + LeakyReLU tip cell per scale
# (layers.py:82-89), then the per-frame head over B*T frames (TimeDistributed, yolo3_temporal.py:468,542-555)
# ------------------------------------------------------------------------------------------------
def make_temporal_sample(rng, windows, T, num_class, size=416):
    """(tips5, tip-cell weights, pred weights, pred biases) for `temporal_head_forward_cpu`."""
    from tests.util import CHANNELS, make_pred_weights, make_tips
    tips5 = make_tips(rng, windows, size=size, T=T)
    tw = [rng.uniform(-0.07, 0.07, (c, c, 3, 1, 1)).astype(np.float32) for c in CHANNELS]
    ws, bs = make_pred_weights(rng, num_class)
    return tips5, tw, ws, bs


def temporal_head_forward_cpu(tips5, tw, ws, bs, num_class, threads=None):
    import torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    mids = []
    with torch.no_grad():
        for x, w in zip(tips5, tw):
            B, T, C, H, W = x.shape
            xt = torch.from_numpy(x).permute(0, 2, 1, 3, 4)                       # swapaxes to (B,C,T,H,W), yolo3_temporal.py:231
            y = torch.nn.functional.conv3d(xt, torch.from_numpy(w), padding=(1, 0, 0))
            y = y / float(np.sqrt(1.0 + 1e-5))                                   # identity BatchNorm (eps 1e-5)
            y = torch.nn.functional.leaky_relu(y, 0.1)
            mids.append(y.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W).contiguous().numpy())
    return head_forward_cpu(mids, ws, bs, num_class, threads=threads)
