"""Oracle: the host-side post-processing loop of detect() (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned by the reference.

Follows detect_yolo3.py:222-261: `bboxes.clip(0, S)` (MXNet fp32 clip, :226), then per image
`valid_pred = where(id >= 0)` (:256), `box / S` (:257, float32 array / python int -> float32), `id.astype(int)` (:258),
rows `[id, score, x1, y1, x2, y2]` (:261-265); the mult_out variant (:238-252) does the same per window offset.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def postprocess(ids, scores, bboxes, size):
    """ids/scores (..., post, 1), bboxes (..., post, 4) -> (rows (F, post, 6) fp32 padded with -1, counts (F,) int32)
    where rows[f, :counts[f]] are the [id, score, x1/S, y1/S, x2/S, y2/S] lists detect() appends for image f, in order."""
    ids = np.asarray(ids, f32)
    post = ids.shape[-2]
    i2 = ids.reshape(-1, post)
    s2 = np.asarray(scores, f32).reshape(-1, post)
    b2 = np.clip(np.asarray(bboxes, f32).reshape(-1, post, 4), f32(0), f32(size)).astype(f32)
    F = i2.shape[0]
    rows = np.full((F, post, 6), -1.0, f32)
    counts = np.zeros((F,), np.int32)
    for f in range(F):
        valid = np.where(i2[f] >= 0)[0]
        n = len(valid)
        counts[f] = n
        rows[f, :n, 0] = i2[f, valid].astype(int).astype(f32)
        rows[f, :n, 1] = s2[f, valid]
        rows[f, :n, 2:] = (b2[f, valid] / f32(size)).astype(f32)
    return rows, counts
