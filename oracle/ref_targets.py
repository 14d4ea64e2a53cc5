"""Oracle: YOLOV3PrefetchTargetGenerator (literal Python double loop on numpy, explicit dtypes).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Own logic pinned against the executed reference source (tests/golden/ref_exec_golden.npz, see oracle/__init__.py); the MXNet operators it calls are restated.

Follows models/definitions/yolo/yolo_target.py:31-148 line by line; the MXNet/GluonCV pieces it
calls (BBoxCornerToCenter, BBoxCenterToCorner, contrib.box_iou, argmax) are restated from
SURVEY.md Appendix A.3/A.4.  The reference ran on NumPy < 2.0 (legacy promotion: np.float32
scalar (op) python int/float -> float64); NumPy here is 2.x, so every promotion is written out.
"""
from __future__ import annotations

import numpy as np

from .ref_nms import box_iou

f32 = np.float32
f64 = np.float64


def corner_to_center(boxes):
    """gluoncv BBoxCornerToCenter(split=True): w=x2-x1; h=y2-y1; x=x1+w/2; y=y1+h/2 (fp32)."""
    b = np.asarray(boxes, dtype=f32)
    x1, y1, x2, y2 = b[..., 0:1], b[..., 1:2], b[..., 2:3], b[..., 3:4]
    w = (x2 - x1).astype(f32)
    h = (y2 - y1).astype(f32)
    x = (x1 + (w / f32(2)).astype(f32)).astype(f32)
    y = (y1 + (h / f32(2)).astype(f32)).astype(f32)
    return x, y, w, h


def prefetch_targets(img_shape, xs_shapes, anchors, offsets, gt_boxes, gt_ids, gt_mixratio=None,
                     num_class=None, return_assign=False, promotion="legacy"):
    """yolo_target.py:31-137.

    img_shape: shape of `img` (only [2],[3] used, :72-73).  xs_shapes: shapes of the 3 fake feature
    maps (only [2],[3] used, :110-111).  anchors: 3 arrays (1,1,3,2) in OUTPUT order (s32,s16,s8).
    offsets: 3 arrays (1,HW_i,1,2) (only their sizes are used, :66).  gt_boxes (B,M,4) corner px,
    padded with -1.  gt_ids (B,M,1) or multi-hot (B,M,C).  gt_mixratio (B,M,1) or None.
    Returns objectness (B,N,1), center (B,N,2), scale (B,N,2), weights (B,N,2), class (B,N,C);
    with return_assign also int32 arrays match (B,M), row (B,M) (final row index, -1 = not written).
    promotion: 'legacy' = the NumPy < 2 scalar promotion of the reference's era (np.float32 scalar (op) python number ->
    float64; the product behaviour); 'nep50' = NumPy >= 2 (stays fp32), used ONLY to compare bit-exactly with golden vectors
    produced by executing the reference's own loop under the NumPy installed here (tests/golden/ref_exec_golden.npz).
    """
    gt_boxes = np.asarray(gt_boxes, dtype=f32)
    gt_ids = np.asarray(gt_ids, dtype=f32)
    all_anchors = np.concatenate([np.asarray(a, dtype=f32).reshape(-1, 2) for a in anchors], 0)  # :62
    num_anchors = np.cumsum([np.asarray(a).size // 2 for a in anchors])                           # :65
    num_offsets = np.cumsum([np.asarray(o).size // 2 for o in offsets])                           # :66
    _offsets = [0] + num_offsets.tolist()                                                         # :67
    orig_height, orig_width = int(img_shape[2]), int(img_shape[3])                                # :72-73
    B, M = gt_boxes.shape[0], gt_boxes.shape[1]
    C = int(num_class) if num_class is not None else (gt_ids.shape[-1] if gt_ids.shape[-1] > 1 else None)
    assert C is not None, "num_class required for single-id labels"
    n_cells, n_anc = _offsets[-1], int(num_anchors[-1])
    center_targets = np.zeros((B, n_cells, n_anc, 2), f32)                                        # :76-79
    scale_targets = np.zeros_like(center_targets)
    weights = np.zeros_like(center_targets)
    objectness = np.zeros((B, n_cells, n_anc, 1), f32)                                            # :81
    class_targets = np.full((B, n_cells, n_anc, C), -1, f32)                                      # :82-83

    gtx, gty, gtw, gth = corner_to_center(gt_boxes)                                               # :88
    shift_gt = np.concatenate((f32(-0.5) * gtw, f32(-0.5) * gth, f32(0.5) * gtw, f32(0.5) * gth), -1).astype(f32)
    half = (all_anchors / f32(2)).astype(f32)                                                     # :90-91
    shift_anchor = np.concatenate((f32(0) * all_anchors - half, f32(0) * all_anchors + half), -1).astype(f32)
    ious = box_iou(shift_anchor, shift_gt).transpose(1, 0, 2)                                     # :92 (B,9,M)
    matches = ious.argmax(axis=1)                                                                 # :94 first max
    valid_gts = (gt_boxes >= 0).prod(axis=-1)                                                     # :95
    single = gt_ids.shape[-1] == 1
    assign_match = np.full((B, M), -1, np.int32)
    assign_row = np.full((B, M), -1, np.int32)

    for b in range(B):                                                                            # :104
        for m in range(M):                                                                        # :105
            if valid_gts[b, m] < 1:                                                               # :106
                break
            match = int(matches[b, m])                                                            # :108
            nlayer = int(np.nonzero(num_anchors > match)[0][0])                                   # :109
            height, width = int(xs_shapes[nlayer][2]), int(xs_shapes[nlayer][3])                  # :110-111
            x, y, w, h = gtx[b, m, 0], gty[b, m, 0], gtw[b, m, 0], gth[b, m, 0]                   # fp32 scalars
            # legacy promotion: np.float32 / python int -> float64
            wide = f64 if promotion == "legacy" else f32
            fx = wide(wide(wide(x) / wide(orig_width)) * wide(width))
            fy = wide(wide(wide(y) / wide(orig_height)) * wide(height))
            loc_x = int(fx)                                                                       # :115
            loc_y = int(fy)                                                                       # :116
            index = _offsets[nlayer] + loc_y * width + loc_x                                      # :118
            center_targets[b, index, match, 0] = f32(wide(fx - wide(loc_x)))                      # :119
            center_targets[b, index, match, 1] = f32(wide(fy - wide(loc_y)))                      # :120
            aw, ah = all_anchors[match, 0], all_anchors[match, 1]
            # python max(gtw, 1): returns the fp32 gtw unless 1 > gtw (then the python int 1)
            sx = f32(np.log(f32(w / aw))) if not (1 > w) else f32(np.log(wide(wide(1) / wide(aw))))   # :121
            sy = f32(np.log(f32(h / ah))) if not (1 > h) else f32(np.log(wide(wide(1) / wide(ah))))   # :122
            scale_targets[b, index, match, 0] = sx
            scale_targets[b, index, match, 1] = sy
            weights[b, index, match, :] = f32(wide(2.0) - wide(wide(wide(f32(w * h)) / wide(orig_width)) / wide(orig_height)))   # :123
            objectness[b, index, match, 0] = gt_mixratio[b, m, 0] if gt_mixratio is not None else 1   # :124-125
            class_targets[b, index, match, :] = 0                                                 # :126
            if single:
                class_targets[b, index, match, int(gt_ids[b, m, 0])] = 1                          # :128
            else:
                class_targets[b, index, match, :] = gt_ids[b, m, :]                               # :130
            assign_match[b, m] = match
            # final (post-_slice) row of this write -- same order as the train-mode predictions
            a_begin = 0 if nlayer == 0 else int(num_anchors[nlayer - 1])
            a_cnt = int(num_anchors[nlayer]) - a_begin
            row_base = sum((_offsets[i + 1] - _offsets[i]) *
                           (int(num_anchors[i]) - (0 if i == 0 else int(num_anchors[i - 1])))
                           for i in range(nlayer))
            # a cell index past the matched layer's own map (loc_y >= height: centre on / below the bottom border) was written
            # into another layer's rows, in anchor columns `_slice` (:139-148) drops: not visible in the outputs -> row -1
            if _offsets[nlayer] <= index < _offsets[nlayer + 1]:
                assign_row[b, m] = row_base + (index - _offsets[nlayer]) * a_cnt + (match - a_begin)
            else:
                assign_match[b, m] = -1

    outs = tuple(_slice(t, num_anchors, num_offsets)
                 for t in (objectness, center_targets, scale_targets, weights, class_targets))  # :132-136
    if return_assign:
        return outs + (assign_match, assign_row)
    return outs


def _slice(x, num_anchors, num_offsets):
    """yolo_target.py:139-148."""
    anchors = [0] + num_anchors.tolist()
    offsets = [0] + num_offsets.tolist()
    ret = []
    for i in range(len(num_anchors)):
        y = x[:, offsets[i]:offsets[i + 1], anchors[i]:anchors[i + 1], :]
        ret.append(y.reshape(y.shape[0], -1, y.shape[-1]))
    return np.concatenate(ret, axis=1)


def default_generator_inputs(size=416, anchors_out_order=None, strides_out_order=(32, 16, 8)):
    """What YOLO3*TrainTransform.__init__ (transforms.py:167-197) extracts from the train-mode net:
    anchors (1,1,3,2) x3, offsets (1,HW,1,2) x3, fake feature-map shapes, in output order."""
    from .ref_head import ANCHORS_OUT_ORDER, make_offsets
    anchors_out_order = ANCHORS_OUT_ORDER if anchors_out_order is None else anchors_out_order
    anchors = [np.asarray(a, f32).reshape(1, 1, -1, 2) for a in anchors_out_order]
    xs_shapes, offsets = [], []
    for s in strides_out_order:
        h = w = size // s
        xs_shapes.append((1, 1, h, w))
        offsets.append(make_offsets()[:, :, :h, :w, :].reshape(1, -1, 1, 2))
    return (1, 3, size, size), xs_shapes, anchors, offsets
