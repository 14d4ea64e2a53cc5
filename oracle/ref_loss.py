"""Oracle: YOLOV3DynamicTargetGeneratorSimple + YOLOV3TargetMerger + YOLOV3Loss (numpy, explicit fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Own logic pinned against the executed reference source (tests/golden/ref_exec_golden.npz, see oracle/__init__.py); the MXNet operators it calls are restated.

Follows:
  * models/definitions/yolo/yolo_target.py:151-205  YOLOV3DynamicTargetGeneratorSimple.hybrid_forward
  * models/definitions/yolo/yolo_target.py:208-281  YOLOV3TargetMerger.hybrid_forward
  * models/definitions/yolo/yolo3.py:507-515        call site (box_preds = concat of the train-mode decode's bbox)
and restates the un-vendored GluonCV pieces they call from their published algorithms (GluonCV 0.4/0.5 era):
  * gluoncv.nn.bbox.BBoxBatchIOU(axis=-1, fmt='corner', offset=0, eps=1e-15)   
  * gluoncv.loss.YOLOV3Loss: objectness / centre sigmoid-BCE, scale L1, class sigmoid-BCE, each a per-sample
    mean over the non-batch axes multiplied back by the element count (i.e. a per-sample sum);
    mxnet.gluon.loss.SigmoidBinaryCrossEntropyLoss(from_sigmoid=False) = relu(x) - x*z + softrelu(-|x|), L1Loss = |z - x|,
    both multiplied by the sample weight before the mean.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def bbox_batch_iou(a, b, offset=0.0, eps=1e-15):
    """BBoxBatchIOU, corner format: a (B,N,4), b (B,M,4) -> (B,N,M), every op in fp32 in GluonCV's order."""
    a = np.asarray(a, f32)
    b = np.asarray(b, f32)
    al, at, ar, ab = [a[..., i][..., :, None] for i in range(4)]
    bl, bt, br, bb = [b[..., i][..., None, :] for i in range(4)]
    left = np.maximum(al, bl)
    right = np.minimum(ar, br)
    top = np.maximum(at, bt)
    bot = np.minimum(ab, bb)
    iw = np.clip((right - left + f32(offset)).astype(f32), f32(0), f32(6.55040e+04))
    ih = np.clip((bot - top + f32(offset)).astype(f32), f32(0), f32(6.55040e+04))
    inter = (iw * ih).astype(f32)
    area_a = ((ar - al + f32(offset)).astype(f32) * (ab - at + f32(offset)).astype(f32)).astype(f32)
    area_b = ((br - bl + f32(offset)).astype(f32) * (bb - bt + f32(offset)).astype(f32)).astype(f32)
    union = ((area_a + area_b).astype(f32) - inter).astype(f32)
    return (inter / (union + f32(eps)).astype(f32)).astype(f32)


def dynamic_targets(box_preds, gt_boxes, num_class, ignore_iou_thresh):
    """yolo_target.py:175-205.  Returns (objness_t, center_t, scale_t, weight_t, class_t)."""
    bp = np.asarray(box_preds, f32).reshape(box_preds.shape[0], -1, 4)
    B, N = bp.shape[:2]
    ious = bbox_batch_iou(bp, gt_boxes)                                   # (B,N,M)
    ious_max = ious.max(axis=-1, keepdims=True)
    objness_t = ((ious_max > f32(ignore_iou_thresh)).astype(f32) * f32(-1)).astype(f32)   # -1 ignored, -0.0 otherwise
    center_t = np.zeros((B, N, 2), f32)
    scale_t = np.zeros((B, N, 2), f32)
    weight_t = np.zeros((B, N, 2), f32)
    class_t = np.full((B, N, num_class), -1.0, f32)
    return objness_t, center_t, scale_t, weight_t, class_t


def target_merge(box_preds, gt_boxes, obj_t, centers_t, scales_t, weights_t, clas_t, num_class, ignore_iou_thresh,
                 label_smooth=False):
    """yolo_target.py:226-281.  Returns [objectness, center_targets, scale_targets, weights, class_targets, class_mask]."""
    dyn = dynamic_targets(box_preds, gt_boxes, num_class, ignore_iou_thresh)
    obj_t = np.asarray(obj_t, f32)
    mask = obj_t > 0                                                       # (B,N,1)
    objectness = np.where(mask, obj_t, dyn[0]).astype(f32)
    mask2 = np.tile(mask, (1, 1, 2))
    center_targets = np.where(mask2, np.asarray(centers_t, f32), dyn[1]).astype(f32)
    scale_targets = np.where(mask2, np.asarray(scales_t, f32), dyn[2]).astype(f32)
    weights = np.where(mask2, np.asarray(weights_t, f32), dyn[3]).astype(f32)
    mask3 = np.tile(mask, (1, 1, num_class))
    class_targets = np.where(mask3, np.asarray(clas_t, f32), dyn[4]).astype(f32)
    smooth_weight = 1.0 / num_class
    if label_smooth:
        smooth_weight = min(1.0 / num_class, 1.0 / 40)
        class_targets = np.where(class_targets > 0.5, (class_targets - f32(smooth_weight)).astype(f32), class_targets)
        class_targets = np.where((class_targets < -0.5) | (class_targets > 0.5), class_targets,
                                 np.full_like(class_targets, f32(smooth_weight))).astype(f32)
    class_mask = (mask3.astype(f32) * (class_targets >= 0).astype(f32)).astype(f32)
    return [objectness, center_targets, scale_targets, weights, class_targets, class_mask]


def _sigmoid_bce(pred, label, weight):
    """SigmoidBinaryCrossEntropyLoss(from_sigmoid=False) elementwise part, fp32 ops (softrelu = log1p(exp(.)))."""
    pred = np.asarray(pred, f32)
    label = np.asarray(label, f32)
    relu = np.maximum(pred, f32(0))
    soft = np.log1p(np.exp(-np.abs(pred)).astype(f32)).astype(f32)
    loss = ((relu - (pred * label).astype(f32)).astype(f32) + soft).astype(f32)
    return (loss * np.asarray(weight, f32)).astype(f32)


def yolo3_loss(objness, box_centers, box_scales, cls_preds, objness_t, center_t, scale_t, weight_t, class_t, class_mask):
    """gluoncv.loss.YOLOV3Loss.hybrid_forward -> (obj_loss, center_loss, scale_loss, cls_loss), each (B,).
    The per-sample means are accumulated in float64 here (the device reduces in fp32 tree order; tests use rtol 1e-4)."""
    objness_t = np.asarray(objness_t, f32)
    B = objness_t.shape[0]
    denorm = f32(np.prod(objness_t.shape[1:]))
    weight_t = (np.asarray(weight_t, f32) * objness_t).astype(f32)
    hard_objness_t = np.where(objness_t > 0, np.ones_like(objness_t), objness_t)
    new_objness_mask = np.where(objness_t > 0, objness_t, (objness_t >= 0).astype(f32)).astype(f32)

    def mean_b(x):
        return x.reshape(B, -1).astype(np.float64).mean(axis=1)

    obj_loss = mean_b(_sigmoid_bce(objness, hard_objness_t, new_objness_mask)) * denorm
    center_loss = mean_b(_sigmoid_bce(box_centers, center_t, weight_t)) * (denorm * 2)
    l1 = (np.abs((np.asarray(scale_t, f32) - np.asarray(box_scales, f32)).astype(f32)) * weight_t).astype(f32)
    scale_loss = mean_b(l1) * (denorm * 2)
    denorm_class = f32(np.prod(np.asarray(class_t).shape[1:]))
    cmask = (np.asarray(class_mask, f32) * objness_t).astype(f32)
    cls_loss = mean_b(_sigmoid_bce(cls_preds, class_t, cmask)) * denorm_class
    return [x.astype(f32) for x in (obj_loss, center_loss, scale_loss, cls_loss)]
