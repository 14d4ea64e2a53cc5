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
