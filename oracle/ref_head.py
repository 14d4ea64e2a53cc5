"""Oracle: 1x1 prediction conv + YOLOOutputV3 decode + scale concat (numpy, fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Own logic pinned against the executed reference source (tests/golden/ref_exec_golden.npz, see oracle/__init__.py); the MXNet operators it calls are restated.

Follows, line by line:
  * models/definitions/yolo/yolo3.py:43-74    (constructor constants: anchors, offsets)
  * models/definitions/yolo/yolo3.py:157-199  (prediction conv, decode, row order)
  * models/definitions/yolo/yolo3.py:522-534  (concat of the three scales, NMS, slice)
  * models/definitions/yolo/wrappers.py:80-84 + yolo3.py:416-417 (anchor / stride order)
Layout is the reference's own: NCHW feature maps, (B, rows, 6) detections.
"""
from __future__ import annotations

import numpy as np

# wrappers.py:80-84 (listed s8, s16, s32) -- yolo3.py:416-417 reverses both lists, so output i=0
# is the stride-32 scale.
ANCHORS_S8_S16_S32 = [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]]
STRIDES_S8_S16_S32 = [8, 16, 32]
ANCHORS_OUT_ORDER = ANCHORS_S8_S16_S32[::-1]      # s32, s16, s8
STRIDES_OUT_ORDER = STRIDES_S8_S16_S32[::-1]      # 32, 16, 8
TIP_CHANNELS_OUT_ORDER = [1024, 512, 256]         # yolo3.py:253 channel*2, channels [512,256,128]

f32 = np.float32


def sigmoid_f32(x: np.ndarray) -> np.ndarray:
    """mshadow_op::sigmoid: 1 / (1 + exp(-x)) in fp32 [MXNet-upstream]."""
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        return (f32(1.0) / (f32(1.0) + np.exp(-x, dtype=f32))).astype(f32)


def conv1x1(x: np.ndarray, weight: np.ndarray, bias: np.ndarray | None) -> np.ndarray:
    """nn.Conv2D(all_pred, kernel_size=1) -- yolo3.py:62,157.  x (B,Cin,H,W), weight (N,Cin,1,1)."""
    x = np.asarray(x, dtype=f32)
    B, Cin, H, W = x.shape
    w2 = np.asarray(weight, dtype=f32).reshape(weight.shape[0], Cin)
    out = np.matmul(w2[None], x.reshape(B, Cin, H * W))          # (B,N,HW) fp32
    if bias is not None:
        out = out + np.asarray(bias, dtype=f32).reshape(1, -1, 1)
    return out.reshape(B, w2.shape[0], H, W).astype(f32)


def make_offsets(alloc_size=(128, 128)) -> np.ndarray:
    """yolo3.py:67-74: offsets[0,0,y,x,:] = (x, y), shape (1,1,Hmax,Wmax,2)."""
    gx, gy = np.meshgrid(np.arange(alloc_size[1]), np.arange(alloc_size[0]))
    off = np.concatenate((gx[:, :, None], gy[:, :, None]), axis=-1)
    return off[None, None].astype(f32)


def decode(pred: np.ndarray, anchors, stride: int, num_class: int, mode: str = "infer",
           alloc_size=(128, 128)):
    """YOLOOutputV3.hybrid_forward after the conv -- yolo3.py:158-199.

    pred: (B, A*(5+C), H, W) fp32.  mode: 'infer' | 'train' | 'agnostic'.
    """
    pred = np.asarray(pred, dtype=f32)
    anchors = np.asarray(anchors, dtype=f32).reshape(1, 1, -1, 2)       # :64
    A = anchors.shape[2]
    P = 5 + num_class
    B, N, H, W = pred.shape
    assert N == A * P, (N, A, P)
    assert H <= alloc_size[0] and W <= alloc_size[1]
    p = pred.reshape(B, A * P, H * W)                                    # :158
    p = p.transpose(0, 2, 1).reshape(B, H * W, A, P)                     # :160
    raw_centers = p[..., 0:2]                                            # :162
    raw_scales = p[..., 2:4]                                             # :163
    objness = p[..., 4:5]                                                # :164
    class_pred = p[..., 5:]                                              # :165
    offsets = make_offsets(alloc_size)[:, :, :H, :W, :].reshape(1, -1, 1, 2)   # :168-170
    centers = ((sigmoid_f32(raw_centers) + offsets).astype(f32) * f32(stride)).astype(f32)  # :172
    with np.errstate(over="ignore"):
        scales = (np.exp(raw_scales, dtype=f32) * anchors).astype(f32)   # :173
    conf = sigmoid_f32(objness)                                          # :174
    class_score = (sigmoid_f32(class_pred) * conf).astype(f32)           # :175
    wh = (scales / f32(2.0)).astype(f32)                                 # :176
    bbox = np.concatenate((centers - wh, centers + wh), axis=-1).astype(f32)   # :177
    if mode == "train":                                                  # :179-182
        return (bbox.reshape(B, -1, 4), raw_centers.copy(), raw_scales.copy(), objness.copy(),
                class_pred.copy(), anchors.copy(), offsets.copy())
    if mode == "agnostic":                                               # :184-188
        ids = (conf * f32(0)).astype(f32) + f32(0)
        det = np.concatenate((ids, conf, bbox), axis=-1)
        return det.reshape(B, -1, 6).astype(f32)
    assert mode == "infer"
    C = num_class
    bboxes = np.tile(bbox[None], (C, 1, 1, 1, 1))                        # :191  (C,B,HW,A,4)
    scores = class_score.transpose(3, 0, 1, 2)[..., None]                # :192  (C,B,HW,A,1)
    ids = (scores * f32(0)).astype(f32) + np.arange(C, dtype=f32).reshape(C, 1, 1, 1, 1)  # :194
    det = np.concatenate((ids, scores, bboxes), axis=-1)                 # :195
    det = det.transpose(1, 0, 2, 3, 4).reshape(B, -1, 6)                 # :197
    return det.astype(f32)


def yolo_output_v3(x, weight, bias, anchors, stride, num_class, mode="infer"):
    """Full block: conv (yolo3.py:157) + decode (:158-199)."""
    return decode(conv1x1(x, weight, bias), anchors, stride, num_class, mode)


def head_detections(tips, weights, biases, num_class, anchors=None, strides=None, mode="infer"):
    """Three scales in output order (s32, s16, s8) concatenated on the row axis -- yolo3.py:523."""
    anchors = ANCHORS_OUT_ORDER if anchors is None else anchors
    strides = STRIDES_OUT_ORDER if strides is None else strides
    dets = [yolo_output_v3(t, w, b, a, s, num_class, mode)
            for t, w, b, a, s in zip(tips, weights, biases, anchors, strides)]
    return np.concatenate(dets, axis=1)


def row_index(num_class, hw_list, scale, c, cell, a, A=3):
    """Closed-form row of detection (scale, class c, cell, anchor a) -- SURVEY A.2."""
    base = sum(num_class * hw * A for hw in hw_list[:scale])
    return base + c * (hw_list[scale] * A) + cell * A + a
