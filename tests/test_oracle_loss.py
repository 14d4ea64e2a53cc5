"""Known-answer tests for the oracle's restatement of BBoxBatchIOU / YOLOV3TargetMerger / YOLOV3Loss
(SURVEY.md 8f row 1).  Values derived by hand from yolo_target.py:151-281 and the GluonCV loss definition."""
import numpy as np

from oracle import ref_loss, ref_nms


def test_bbox_batch_iou_matches_box_iou_and_doc_example():
    a = np.array([[[0.5, 0.5, 1, 1], [0, 0, 0.5, 0.5]]], np.float32)
    b = np.array([[[0.25, 0.25, 0.75, 0.75]]], np.float32)
    iou = ref_loss.bbox_batch_iou(a, b)
    np.testing.assert_allclose(iou[0, :, 0], [0.0625 / 0.4375, 0.0625 / 0.4375], rtol=1e-6)
    rng = np.random.RandomState(0)
    xy = rng.uniform(0, 300, (1, 40, 2)); wh = rng.uniform(5, 120, (1, 40, 2))
    a = np.concatenate([xy, xy + wh], -1).astype(np.float32)
    xy = rng.uniform(0, 300, (1, 7, 2)); wh = rng.uniform(5, 120, (1, 7, 2))
    b = np.concatenate([xy, xy + wh], -1).astype(np.float32)
    ref = ref_nms.box_iou(a[0], b[0])                       # MXNet contrib.box_iou restatement (no eps)
    np.testing.assert_allclose(ref_loss.bbox_batch_iou(a, b)[0], ref, rtol=1e-5, atol=1e-7)


def test_padding_gt_gives_zero_iou_and_ignore_rule_is_strict():
    box = np.array([[[10, 10, 50, 50], [100, 100, 140, 140], [10, 10, 50, 30]]], np.float32)       # (1,3,4)
    gt = np.array([[[10, 10, 50, 50], [-1, -1, -1, -1]]], np.float32)
    obj, ctr, scl, wgt, cls = ref_loss.dynamic_targets(box, gt, num_class=3, ignore_iou_thresh=0.5)
    # IoUs with the real GT: 1.0, 0.0, 0.5 (exactly 0.5 is NOT > 0.5); padding GT contributes 0
    assert obj[0, :, 0].tolist() == [-1.0, 0.0, 0.0]
    assert np.all(ctr == 0) and np.all(scl == 0) and np.all(wgt == 0) and np.all(cls == -1)


def test_merger_prefetched_positive_overrides_ignore():
    box = np.array([[[10, 10, 50, 50], [100, 100, 140, 140]]], np.float32)
    gt = np.array([[[10, 10, 50, 50]]], np.float32)
    obj_t = np.array([[[1.0], [0.0]]], np.float32)          # anchor 0 is the prefetched positive
    ctr_t = np.array([[[0.3, 0.7], [0.9, 0.9]]], np.float32)
    scl_t = np.array([[[0.1, -0.2], [5, 5]]], np.float32)
    wgt_t = np.array([[[1.5, 1.5], [9, 9]]], np.float32)
    cls_t = np.array([[[0, 1, 0], [1, 1, 1]]], np.float32)
    o, c, s, w, k, m = ref_loss.target_merge(box, gt, obj_t, ctr_t, scl_t, wgt_t, cls_t, 3, 0.7)
    assert o[0, :, 0].tolist() == [1.0, 0.0]                # positive wins over the -1 it would get dynamically
    assert c[0].tolist() == [[np.float32(0.3), np.float32(0.7)], [0.0, 0.0]]
    assert w[0].tolist() == [[1.5, 1.5], [0.0, 0.0]]
    assert k[0].tolist() == [[0, 1, 0], [-1, -1, -1]]
    assert m[0].tolist() == [[1, 1, 1], [0, 0, 0]]
    # label smoothing: 1 -> 1 - 1/40 (C=3 -> min(1/3, 1/40)), 0 -> 1/40, -1 stays
    o2, _, _, _, k2, m2 = ref_loss.target_merge(box, gt, obj_t, ctr_t, scl_t, wgt_t, cls_t, 3, 0.7, label_smooth=True)
    np.testing.assert_allclose(k2[0, 0], [1 / 40, 1 - 1 / 40, 1 / 40], rtol=1e-6)
    assert k2[0, 1].tolist() == [-1, -1, -1]


def test_loss_hand_values():
    # one sample, two anchors, one class; anchor 0 positive (objness_t 1), anchor 1 ignored (-1)
    objness = np.array([[[0.0], [3.0]]], np.float32)
    centers = np.zeros((1, 2, 2), np.float32); scales = np.array([[[0.5, -0.5], [1, 1]]], np.float32)
    cls_p = np.zeros((1, 2, 1), np.float32)
    obj_t = np.array([[[1.0], [-1.0]]], np.float32)
    ctr_t = np.full((1, 2, 2), 0.5, np.float32); scl_t = np.zeros((1, 2, 2), np.float32)
    wgt_t = np.array([[[2.0, 2.0], [0, 0]]], np.float32)
    cls_t = np.array([[[1.0], [-1.0]]], np.float32); cls_m = np.array([[[1.0], [0.0]]], np.float32)
    lo, lc, ls, lk = ref_loss.yolo3_loss(objness, centers, scales, cls_p, obj_t, ctr_t, scl_t, wgt_t, cls_t, cls_m)
    ln2 = np.log(2.0)
    np.testing.assert_allclose(lo, [ln2], rtol=1e-6)        # anchor 0: bce(0,1)=ln2, mask 1; anchor 1 masked out
    np.testing.assert_allclose(lc, [2 * 2 * ln2], rtol=1e-6)   # two coords, bce(0, .5) = ln2, weight 2*1
    np.testing.assert_allclose(ls, [2 * (0.5 + 0.5)], rtol=1e-6)
    np.testing.assert_allclose(lk, [ln2], rtol=1e-6)
