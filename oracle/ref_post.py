"""Oracle: the host-side post-processing loop of detect() (numpy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  `hierarchical_nms` is pinned bit-exactly by golden vectors produced by executing
the reference's own function (tests/golden/hier_nms_golden.npz); `postprocess` (numpy restatement of detect()'s inline loop) is unpinned.

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


def _iou_pixel(bb, bbgt):
    """detect_yolo3.py:712-733 `iou`: PASCAL-style `+ 1` extents, plain float64 arithmetic in this operation order."""
    ov = 0.0
    iw = min(bb[2], bbgt[2]) - max(bb[0], bbgt[0]) + 1
    ih = min(bb[3], bbgt[3]) - max(bb[1], bbgt[1]) + 1
    if iw > 0 and ih > 0:
        intersect = iw * ih
        ua = (bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) + (bbgt[2] - bbgt[0] + 1.) * (bbgt[3] - bbgt[1] + 1.) - intersect
        ov = intersect / ua
    return ov


def hierarchical_nms(rows, counts, levels, parent, branch, ov_thresh=0.5, conf_thresh=0.0, level_thresh=10, stats=None):
    """detect_yolo3.py:736-789 on packed rows.  rows (F, post, 6) [cls, conf, x1, y1, x2, y2] with counts (F,) valid rows per
    image (the `predictions[img]` lists, in order); levels (C,) = dataset.get_levels() (combined.py:117-126); parent (C,) =
    class index of the parent (-1 under ROOT), i.e. `cls_map.index(parents[cls_map[cls]])` (:766); branch (C,C) =
    dataset.on_branch(i, j) (combined.py:143-150, :742-746).  Arithmetic: float64 on the float32-representable inputs = what the
    reference computes on predictions re-loaded from its .txt files (Python floats).  Returns (out_rows (F,post,6) fp32 padded
    with -1, out_counts (F,) int32): the `new_predictions[img]` lists in order."""
    rows = np.asarray(rows, f32)
    F, post, _ = rows.shape
    level_thresh = max(0, level_thresh)                                       # :749
    out = np.full((F, post, 6), -1.0, f32)
    ocnt = np.zeros((F,), np.int32)
    for f in range(F):
        boxes = [[int(r[0]), float(r[1])] + [float(v) for v in r[2:]] for r in rows[f, :counts[f]]]
        new = []
        for box in sorted(boxes, key=lambda x: x[0], reverse=True):           # :756 stable, highest (most leafy) class first
            cls, conf, coords = box[0], box[1], box[2:]
            if conf < conf_thresh:                                            # :761
                continue
            while levels[cls] > level_thresh:                                 # :765-766
                cls = int(parent[cls])
            max_ov, max_idx = 0, -1
            for idx, boxb in enumerate(new):                                  # :771-775 first strictly-largest overlap above the threshold
                overlap = _iou_pixel(coords, boxb[2:])
                if overlap > ov_thresh and overlap > max_ov:
                    max_ov, max_idx = overlap, idx
            if max_idx == -1:                                                 # :777-778
                new.append([cls, conf] + coords)
                if stats is not None: stats["new"] = stats.get("new", 0) + 1
            else:
                boxb = new[max_idx]
                if not branch[cls][boxb[0]]:                                  # :782-783
                    new.append([cls, conf] + coords)
                    if stats is not None: stats["off_branch"] = stats.get("off_branch", 0) + 1
                elif cls == boxb[0]:                                          # :785-786
                    new[max_idx][1] = max(new[max_idx][1], conf)
                    if stats is not None: stats["same_cls_max"] = stats.get("same_cls_max", 0) + 1
                elif stats is not None:                                       # :787 ignored: a child already stands there
                    stats["ignored"] = stats.get("ignored", 0) + 1
        ocnt[f] = len(new)
        for j, b in enumerate(new):
            out[f, j] = b
    return out, ocnt
