"""Known-answer cases for the assumptions the box_nms / box_iou restatement makes about the UPSTREAM MXNet operator
(oracle/ASSUMPTIONS.md lists them; ids A1..A12 below refer to that file).  Each case: (name, data (N,6) rows
[id, score, x1, y1, x2, y2], box_nms kwargs, expected `record` = original row of every kept element in output order,
-1 padding).  The expected records are derived BY HAND from the stated assumption, not produced by any implementation;
the CPU suite runs them on the oracle (C and numpy twins), the GPU suite on vd_box_nms."""
import numpy as np

f32 = np.float32
NAN = float("nan")


def _r(i, s, x1, y1, x2, y2):
    return [i, s, x1, y1, x2, y2]


CASES = [
    # A1: strict `score > valid_thresh`; NaN is never valid
    ("A1_strict_valid_and_nan",
     [_r(0, 0.01, 0, 0, 10, 10), _r(0, 0.02, 50, 50, 60, 60), _r(0, NAN, 100, 100, 110, 110)],
     dict(overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0), [1, -1, -1]),
    # A2: background_id drops rows whose (int) id equals it -- only when id_index >= 0 and background_id >= 0
    ("A2_background_id",
     [_r(0, 0.9, 0, 0, 10, 10), _r(1, 0.8, 50, 50, 60, 60), _r(1.7, 0.7, 100, 100, 110, 110), _r(2, 0.6, 150, 150, 160, 160)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0, background_id=1), [0, 3, -1, -1]),
    ("A2_background_id_ignored_without_id_index",
     [_r(0, 0.9, 0, 0, 10, 10), _r(1, 0.8, 50, 50, 60, 60)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=-1, background_id=1), [0, 1]),
    # A3: BoxArea() is 0 for a box with a negative extent.  Row 1 is inverted in x AND y: its signed product would be +100
    # (like a real 10x10 box).  inter(0,1) = 0 either way -> kept; the case pins that a degenerate box neither suppresses nor
    # is suppressed, and A3b pins the value of the area through a 3-box chain.
    ("A3_degenerate_box_never_overlaps",
     [_r(0, 0.9, 0, 0, 10, 10), _r(0, 0.8, 10, 10, 0, 0), _r(0, 0.7, 1, 1, 9, 9)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0), [0, 1, -1]),
    # A4: stable descending sort -- equal scores keep ascending original order
    ("A4_ties_by_row",
     [_r(0, 0.5, 200, 200, 210, 210), _r(0, 0.7, 0, 0, 10, 10), _r(0, 0.5, 100, 100, 110, 110), _r(0, 0.5, 300, 300, 310, 310)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0), [1, 0, 2, 3]),
    # A5: topk <= 0 means "all"; the cut happens BEFORE suppression (rank 2 is dropped although rank 1 dies)
    ("A5_topk_cut_before_suppression",
     [_r(0, 0.9, 0, 0, 10, 10), _r(0, 0.8, 1, 0, 11, 10), _r(0, 0.7, 100, 100, 110, 110)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=2, id_index=0), [0, -1, -1]),
    ("A5_topk_zero_is_all",
     [_r(0, 0.9, 0, 0, 10, 10), _r(0, 0.8, 1, 0, 11, 10), _r(0, 0.7, 100, 100, 110, 110)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=0, id_index=0), [0, 2, -1]),
    # A6: class ids are compared after truncation to int: 1.2 and 1.9 are the same class, 1.9 and 2.0 are not
    ("A6_ids_truncate_to_int",
     [_r(1.2, 0.9, 0, 0, 10, 10), _r(1.9, 0.8, 0, 0, 10, 10), _r(2.0, 0.7, 0, 0, 10, 10)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0), [0, 2, -1]),
    ("A6_force_suppress_ignores_ids",
     [_r(1, 0.9, 0, 0, 10, 10), _r(2, 0.8, 0, 0, 10, 10)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0, force_suppress=True), [0, -1]),
    # A7: suppression needs IoU strictly greater than the threshold: IoU(0,1) = 50/150 = 1/3 exactly representable? no ->
    # use IoU = 0.5 exactly: boxes 10x10 and the same shifted by 1/3 of ... simpler: [0,0,10,10] vs [0,0,10,5]: IoU = 50/100 = 0.5
    ("A7_strict_iou_gt_thresh",
     [_r(0, 0.9, 0, 0, 10, 10), _r(0, 0.8, 0, 0, 10, 5)],
     dict(overlap_thresh=0.5, valid_thresh=0.0, topk=-1, id_index=0), [0, 1]),
    ("A7_iou_above_thresh",
     [_r(0, 0.9, 0, 0, 10, 10), _r(0, 0.8, 0, 0, 10, 5)],
     dict(overlap_thresh=0.49, valid_thresh=0.0, topk=-1, id_index=0), [0, -1]),
    # A8: center format: extents are x -/+ w/2; same geometry as A7_iou_above_thresh written as (cx, cy, w, h)
    ("A8_center_format",
     [_r(0, 0.9, 5, 5, 10, 10), _r(0, 0.8, 5, 2.5, 10, 5), _r(0, 0.7, 50, 50, 10, 10)],
     dict(overlap_thresh=0.49, valid_thresh=0.0, topk=-1, id_index=0, in_format="center", out_format="center"), [0, 2, -1]),
    # A9: a suppressed box does not suppress (chain A>B>C with IoU(A,C) below the threshold)
    ("A9_chain",
     [_r(0, 0.9, 0, 0, 10, 10), _r(0, 0.8, 2, 0, 12, 10), _r(0, 0.7, 4, 0, 14, 10)],
     dict(overlap_thresh=0.45, valid_thresh=0.0, topk=-1, id_index=0), [0, 2, -1]),
    # A10: nothing valid -> every output row is -1
    ("A10_all_filtered",
     [_r(0, 0.3, 0, 0, 10, 10), _r(0, 0.2, 20, 20, 30, 30)],
     dict(overlap_thresh=0.45, valid_thresh=0.5, topk=-1, id_index=0), [-1, -1]),
]


def case_arrays():
    for name, rows, kw, rec in CASES:
        yield name, np.array(rows, f32), kw, np.array(rec, np.int32)


def expected_output(data, kw, rec):
    """Output tensor implied by `rec`: kept rows copied unmodified in order (A9/A10), -1 elsewhere."""
    out = np.full_like(data, -1.0)
    for pos, r in enumerate(rec):
        if r >= 0:
            out[pos] = data[r]
    return out
