"""Known-answer tests pinning the oracle (CPU).  The reference ships no tests,
so these are hand-derived from Appendix A plus the two upstream operator-doc examples."""
import numpy as np
import pytest

from oracle import ref_head, ref_nms, ref_targets, ref_temporal
from tests.util import random_dets

f32 = np.float32


# ------------------------------------------------------------------ box_nms / box_iou (A.3)
X_DOC = np.array([[0, 0.5, 0.1, 0.1, 0.2, 0.2], [1, 0.4, 0.1, 0.1, 0.2, 0.2],
                  [0, 0.3, 0.1, 0.1, 0.14, 0.14], [2, 0.6, 0.5, 0.5, 0.7, 0.8]], f32)


def test_kat0_box_nms_doc_example_force():
    out = ref_nms.box_nms(X_DOC, overlap_thresh=0.1, coord_start=2, score_index=1, id_index=0, force_suppress=True)
    exp = np.array([[2, 0.6, 0.5, 0.5, 0.7, 0.8], [0, 0.5, 0.1, 0.1, 0.2, 0.2], [-1] * 6, [-1] * 6], f32)
    np.testing.assert_array_equal(out, exp)


def test_kat0_box_nms_doc_example_class_aware():
    out, rec = ref_nms.box_nms(X_DOC, overlap_thresh=0.1, coord_start=2, score_index=1, id_index=0,
                               force_suppress=False, return_record=True)
    exp = np.array([[2, 0.6, 0.5, 0.5, 0.7, 0.8], [0, 0.5, 0.1, 0.1, 0.2, 0.2], [1, 0.4, 0.1, 0.1, 0.2, 0.2], [-1] * 6], f32)
    np.testing.assert_array_equal(out, exp)
    np.testing.assert_array_equal(rec, [3, 0, 1, -1])


def test_kat1_box_iou_doc_example():
    iou = ref_nms.box_iou(np.array([[0.5, 0.5, 1, 1], [0, 0, 0.5, 0.5]], f32), np.array([[0.25, 0.25, 0.75, 0.75]], f32))
    np.testing.assert_allclose(iou, [[0.0625 / 0.4375], [0.0625 / 0.4375]], rtol=1e-6)


def _row(i, s, x1, y1, x2, y2):
    return [i, s, x1, y1, x2, y2]


def test_nms_strict_valid_thresh_and_ties():
    d = np.array([_row(0, 0.01, 0, 0, 10, 10),       # == valid_thresh -> dropped (strict >)
                  _row(0, 0.5, 100, 100, 110, 110),  # tie with next row -> lower row first
                  _row(0, 0.5, 200, 200, 210, 210),
                  _row(0, np.nan, 0, 0, 1, 1)], f32)  # NaN invalid
    out, rec = ref_nms.box_nms(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, return_record=True)
    np.testing.assert_array_equal(rec, [1, 2, -1, -1])
    np.testing.assert_array_equal(out[2:], -1)


def test_nms_topk_drops_valid_rank():
    d = np.array([_row(0, 0.9, 0, 0, 10, 10), _row(0, 0.8, 100, 100, 110, 110), _row(0, 0.7, 200, 200, 210, 210)], f32)
    _, rec = ref_nms.box_nms(d, overlap_thresh=0.45, valid_thresh=0.01, topk=2, id_index=0, return_record=True)
    np.testing.assert_array_equal(rec, [0, 1, -1])


def test_nms_chain_suppressed_box_does_not_suppress():
    # A overlaps B (IoU .667 > .45), B overlaps C (.667), A vs C (.333) -> B dies, C survives
    d = np.array([_row(0, 0.9, 0, 0, 10, 10), _row(0, 0.8, 2, 0, 12, 10), _row(0, 0.7, 4, 0, 14, 10)], f32)
    _, rec = ref_nms.box_nms(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, return_record=True)
    np.testing.assert_array_equal(rec, [0, 2, -1])


def test_nms_class_aware_vs_force_and_all_filtered_and_4d():
    d = np.array([_row(0, 0.9, 0, 0, 10, 10), _row(1, 0.8, 0, 0, 10, 10)], f32)
    _, rec = ref_nms.box_nms(d, 0.45, 0.01, 400, id_index=0, return_record=True)
    np.testing.assert_array_equal(rec, [0, 1])
    _, rec = ref_nms.box_nms(d, 0.45, 0.01, 400, id_index=0, force_suppress=True, return_record=True)
    np.testing.assert_array_equal(rec, [0, -1])
    out = ref_nms.box_nms(d, 0.45, 0.95, 400, id_index=0)
    np.testing.assert_array_equal(out, -1)
    d4 = np.stack([d, d[::-1]])[None]                 # (1,2,2,6): leading dims flatten to batch
    out4, rec4 = ref_nms.box_nms(d4, 0.45, 0.01, 400, id_index=0, return_record=True)
    assert out4.shape == d4.shape
    np.testing.assert_array_equal(rec4[0], [[0, 1], [1, 0]])


def test_upstream_assumption_kats():
    """oracle/ASSUMPTIONS.md: every assumption about the upstream operator, hand-derived expected keep records."""
    from tests.kat_nms import case_arrays, expected_output
    for name, d, kw, rec in case_arrays():
        out, got = ref_nms.box_nms(d, return_record=True, **kw)
        np.testing.assert_array_equal(got, rec, err_msg=name)
        if kw.get("in_format", "corner") == kw.get("out_format", "corner"):
            np.testing.assert_array_equal(out, expected_output(d, kw, rec), err_msg=name)
        if kw.get("in_format", "corner") == "corner":               # the numpy twin restates the corner format only
            kw2 = {k: v for k, v in kw.items() if k not in ("in_format", "out_format")}
            _, got2 = ref_nms.box_nms_py(d, **kw2)
            np.testing.assert_array_equal(got2, rec, err_msg=name + " (numpy twin)")


def test_box_area_clamp_value():
    """A3: the area of an inverted box is 0, not the (positive) product of its two negative extents.  B = [4,0,14,10] (area 100)
    overlaps the inverted row A' = [10,10,0,0] nowhere (Intersect clamps), but overlaps C = [0,0,10,10] with inter 60:
    IoU(C,B) = 60 / (100 + 100 - 60) = 0.4286.  box_iou of the inverted box with anything is exactly 0 (A11)."""
    iou = ref_nms.box_iou(np.array([[10, 10, 0, 0]], f32), np.array([[0, 0, 10, 10], [2, 2, 8, 8]], f32))
    np.testing.assert_array_equal(iou, [[0, 0]])
    iou = ref_nms.box_iou(np.array([[0, 0, 10, 10]], f32), np.array([[4, 0, 14, 10]], f32))
    np.testing.assert_allclose(iou, [[60 / 140]], rtol=1e-6)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("force", [False, True])
def test_c_oracle_matches_python_restatement(seed, force):
    rng = np.random.RandomState(seed)
    d = random_dets(rng, 3, 150, num_class=3, tie_frac=0.3)
    o1, r1 = ref_nms.box_nms(d, 0.45, 0.01, 60, id_index=0, force_suppress=force, return_record=True)
    o2, r2 = ref_nms.box_nms_py(d, 0.45, 0.01, 60, id_index=0, force_suppress=force)
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(o1, o2)


def test_nms_center_format_matches_corner():
    rng = np.random.RandomState(5)
    d = random_dets(rng, 2, 80, num_class=2)
    c = d.copy()
    c[..., 2] = (d[..., 2] + d[..., 4]) / 2; c[..., 3] = (d[..., 3] + d[..., 5]) / 2
    c[..., 4] = d[..., 4] - d[..., 2]; c[..., 5] = d[..., 5] - d[..., 3]
    _, r1 = ref_nms.box_nms(d, 0.45, 0.01, 50, id_index=0, return_record=True)
    out_c, r2 = ref_nms.box_nms(c, 0.45, 0.01, 50, id_index=0, in_format="center", out_format="corner", return_record=True)
    assert (r1 == r2).mean() > 0.97            # same geometry up to fp32 rounding of the conversion
    k = (r2[0] >= 0).sum()
    np.testing.assert_allclose(out_c[0, :k, 2:], d[0][r2[0, :k], 2:], rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------ decode (A.2)
def test_decode_zero_logits_kat():
    C, H, W, stride = 4, 3, 5, 16
    anchors = [30, 61, 62, 45, 59, 119]
    pred = np.zeros((2, 3 * (5 + C), H, W), f32)
    det = ref_head.decode(pred, anchors, stride, C)
    assert det.shape == (2, C * H * W * 3, 6)
    np.testing.assert_allclose(det[..., 1], 0.25, rtol=1e-6)         # sigmoid(0)^2
    for c in range(C):
        for cell in range(H * W):
            for a in range(3):
                row = c * H * W * 3 + cell * 3 + a
                x, y = cell % W, cell // W
                cx, cy = (x + 0.5) * stride, (y + 0.5) * stride
                aw, ah = anchors[2 * a], anchors[2 * a + 1]
                np.testing.assert_allclose(det[0, row], [c, 0.25, cx - aw / 2, cy - ah / 2, cx + aw / 2, cy + ah / 2], rtol=1e-6)


def test_decode_row_order_and_modes():
    rng = np.random.RandomState(3)
    C, H, W = 3, 2, 4
    pred = rng.standard_normal((1, 3 * (5 + C), H, W)).astype(f32)
    anchors = [10, 13, 16, 30, 33, 23]
    det = ref_head.decode(pred, anchors, 8, C)
    bbox, rc, rs, ob, cp, anc, off = ref_head.decode(pred, anchors, 8, C, mode="train")
    assert bbox.shape == (1, H * W * 3, 4) and rc.shape == (1, H * W, 3, 2) and cp.shape == (1, H * W, 3, C)
    P = 5 + C
    for c in range(C):
        for cell in range(H * W):
            for a in range(3):
                row = c * H * W * 3 + cell * 3 + a
                y, x = divmod(cell, W)
                to, tc = pred[0, a * P + 4, y, x], pred[0, a * P + 5 + c, y, x]
                s = ref_head.sigmoid_f32(tc) * ref_head.sigmoid_f32(to)
                np.testing.assert_allclose(det[0, row, 1], s, rtol=1e-6)
                assert det[0, row, 0] == c
                np.testing.assert_array_equal(det[0, row, 2:], bbox[0, cell * 3 + a])
                assert cp[0, cell, a, c] == tc and ob[0, cell, a, 0] == to
    np.testing.assert_array_equal(off[0, :, 0, 0], np.tile(np.arange(W), H))
    ag = ref_head.decode(pred, anchors, 8, C, mode="agnostic")
    assert ag.shape == (1, H * W * 3, 6)
    np.testing.assert_array_equal(ag[0, :, 0], 0)
    np.testing.assert_array_equal(ag[0, :, 2:], bbox[0])
    hw = [4, 16, 64]
    assert ref_head.row_index(20, hw, 2, 3, 5, 1) == 20 * (4 + 16) * 3 + 3 * 64 * 3 + 5 * 3 + 1


def test_head_concat_order_s32_first():
    rng = np.random.RandomState(0)
    C = 2
    tips = [rng.standard_normal((1, 8, h, h)).astype(f32) for h in (1, 2, 4)]
    ws = [rng.uniform(-.1, .1, (3 * (5 + C), 8, 1, 1)).astype(f32) for _ in range(3)]
    bs = [np.zeros(3 * (5 + C), f32)] * 3
    det = ref_head.head_detections(tips, ws, bs, C)
    assert det.shape == (1, C * 3 * (1 + 4 + 16), 6)
    d0 = ref_head.yolo_output_v3(tips[0], ws[0], bs[0], ref_head.ANCHORS_OUT_ORDER[0], 32, C)
    np.testing.assert_array_equal(det[:, :d0.shape[1]], d0)
    assert ref_head.ANCHORS_OUT_ORDER[0] == [116, 90, 156, 198, 373, 326] and ref_head.STRIDES_OUT_ORDER == [32, 16, 8]


# ------------------------------------------------------------------ targets (A.4)
def _gen(gt, ids, mix=None, C=20, size=416, **kw):
    img, xs, anchors, offsets = ref_targets.default_generator_inputs(size)
    return ref_targets.prefetch_targets(img, xs, anchors, offsets, gt, ids, mix, num_class=C, return_assign=True, **kw)


def test_targets_single_gt_kat():
    gt = np.full((1, 3, 4), -1, f32); ids = np.full((1, 3, 1), -1, f32)
    gt[0, 0] = [100, 120, 220, 300]; ids[0, 0, 0] = 7          # w=120,h=180,cx=160,cy=210
    obj, ctr, scl, wgt, cls, match, row = _gen(gt, ids)
    # best zero-centred IoU: anchor (156,198) -> index 1 -> layer 0 (13x13)
    assert match[0, 0] == 1
    fx, fy = 160 / 416 * 13, 210 / 416 * 13
    lx, ly = int(fx), int(fy)
    r = (ly * 13 + lx) * 3 + 1
    assert row[0, 0] == r and (row[0, 1:] == -1).all()
    assert obj.shape == (1, 10647, 1) and cls.shape == (1, 10647, 20)
    np.testing.assert_allclose(ctr[0, r], [fx - lx, fy - ly], rtol=1e-6)
    np.testing.assert_allclose(scl[0, r], [np.log(120 / 156), np.log(180 / 198)], rtol=1e-5)
    np.testing.assert_allclose(wgt[0, r], 2 - 120 * 180 / 416 / 416, rtol=1e-6)
    assert obj[0, r, 0] == 1 and obj.sum() == 1
    exp = np.zeros(20, f32); exp[7] = 1
    np.testing.assert_array_equal(cls[0, r], exp)
    assert (np.delete(cls[0], r, axis=0) == -1).all()


def test_targets_last_writer_wins_break_multihot_mix():
    C = 6
    gt = np.full((1, 5, 4), -1, f32)
    gt[0, 0] = [100, 120, 220, 300]
    gt[0, 1] = [101, 121, 221, 301]          # same cell + anchor -> overwrites GT 0
    gt[0, 2] = [-1, 10, 50, 60]              # invalid -> BREAK
    gt[0, 3] = [10, 10, 40, 50]              # never processed
    ids = np.zeros((1, 5, C), f32); ids[0, 0, [0, 1]] = 1; ids[0, 1, [2]] = 1; ids[0, 3, [5]] = 1
    mix = np.full((1, 5, 1), 0.25, f32); mix[0, 1, 0] = 0.75
    obj, ctr, scl, wgt, cls, match, row = _gen(gt, ids, mix, C=C)
    assert row[0, 0] == row[0, 1] >= 0 and row[0, 2] == -1 and row[0, 3] == -1
    r = row[0, 1]
    np.testing.assert_array_equal(cls[0, r], ids[0, 1])          # replaced, not OR-ed
    assert obj[0, r, 0] == np.float32(0.75) and np.count_nonzero(obj) == 1


def test_targets_slice_layout_row_formula():
    gt = np.full((1, 2, 4), -1, f32); ids = np.zeros((1, 2, 1), f32)
    gt[0, 0] = [200, 200, 212, 216]          # tiny box -> anchor (10,13) = index 6 -> layer 2 (52x52)
    obj, ctr, scl, wgt, cls, match, row = _gen(gt, ids)
    assert match[0, 0] == 6
    cx, cy = 206, 208
    cell = int(cy / 416 * 52) * 52 + int(cx / 416 * 52)
    assert row[0, 0] == 3 * (169 + 676) + cell * 3 + 0
    assert obj[0, row[0, 0], 0] == 1


def test_targets_small_box_float64_log_path():
    gt = np.full((1, 1, 4), -1, f32); ids = np.zeros((1, 1, 1), f32)
    gt[0, 0] = [50, 60, 50.5, 60.25]         # w,h < 1 -> max(gtw,1) returns python int 1 -> float64 log
    obj, ctr, scl, wgt, cls, match, row = _gen(gt, ids)
    aw, ah = 10, 13
    assert match[0, 0] == 6
    np.testing.assert_allclose(scl[0, row[0, 0]], [np.log(1 / aw), np.log(1 / ah)], rtol=1e-6)


def test_targets_centre_on_bottom_border_is_sliced_away():
    """A GT whose centre row maps to loc_y == H of its layer (gy == orig_h) is written at cell index >= HW, i.e. into the NEXT
    layer's rows of the (B, sum HW, 9, .) scratch but in its own anchor columns, which `_slice` drops (yolo_target.py:139-148):
    no visible positive.  (the device once wrote a bogus positive into the next scale.)"""
    gt = np.full((1, 3, 4), -1, f32); ids = np.zeros((1, 3, 1), f32)
    gt[0, 0] = [100, 371, 220, 461]          # w=120,h=90 -> anchor (116,90) = index 0 -> layer 0; cy = 416 -> loc_y = 13
    gt[0, 1] = [100, 120, 220, 300]          # an ordinary GT after it is still processed (no break: the box is valid)
    obj, ctr, scl, wgt, cls, match, row = _gen(gt, ids)
    assert row[0, 0] == -1 and match[0, 0] == -1 and row[0, 1] >= 0
    assert np.count_nonzero(obj) == 1 and obj[0, row[0, 1], 0] == 1
    assert (np.delete(cls[0], row[0, 1], axis=0) == -1).all()


# ------------------------------------------------------------------ temporal (A.5)
def test_temporal_conv_zero_padding_and_identity_bn():
    rng = np.random.RandomState(0)
    B, T, C, H, W = 1, 5, 4, 2, 2
    x = rng.standard_normal((B, T, C, H, W)).astype(f32)
    w = np.zeros((C, C, 3), f32)
    w[:, :, 0] = np.eye(C)                   # y[t] = x[t-1]
    y = ref_temporal.temporal_conv_bn_lrelu(x, w, np.ones(C), np.zeros(C), np.zeros(C), np.ones(C) - 1e-5)
    exp = np.concatenate([np.zeros_like(x[:, :1]), x[:, :-1]], 1)
    exp = np.where(exp > 0, exp, 0.1 * exp)
    np.testing.assert_allclose(y, exp, rtol=1e-6, atol=1e-7)


def test_time_distributed_pool_cat():
    x = np.arange(2 * 3 * 4 * 2 * 2, dtype=f32).reshape(2, 3, 4, 2, 2)
    y = ref_temporal.time_distributed(lambda z: z * 2, x)
    np.testing.assert_array_equal(y, x * 2)
    np.testing.assert_array_equal(ref_temporal.temporal_pooling(x, "max"), x[:, 2])
    np.testing.assert_allclose(ref_temporal.temporal_pooling(x, "mean"), x.mean(1))
    np.testing.assert_array_equal(ref_temporal.late_cat(x)[:, 4:8], x[:, 1])
