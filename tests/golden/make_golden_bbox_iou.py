"""Generates tests/golden/bbox_iou_golden.npz by importing the reference's own pure-numpy helper
`utils/bbox.py::bbox_iou` (the only hot-path-adjacent arithmetic in /root/reference that imports
without MXNet).  Run in the authoring container only; the .npz travels to the GPU box."""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/utils/bbox.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bbox_iou_golden.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_utils_bbox", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.RandomState(20240611)
    xy = rng.uniform(0, 400, size=(64, 2))
    wh = np.exp(rng.uniform(np.log(4), np.log(300), size=(64, 2)))
    a = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    xy = rng.uniform(0, 400, size=(48, 2))
    wh = np.exp(rng.uniform(np.log(4), np.log(300), size=(48, 2)))
    b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    b[:8] = a[:8]                                  # identical boxes -> IoU 1
    b[8:12, :2] = a[8:12, 2:4] + 1.0               # disjoint boxes -> IoU 0
    b[8:12, 2:] = b[8:12, :2] + 5.0
    iou = mod.bbox_iou(a.astype(np.float64), b.astype(np.float64))   # float64 ground truth
    iou32 = mod.bbox_iou(a, b)                                       # the helper on fp32 inputs
    # zero-centred anchors vs zero-centred GTs, as yolo_target.py:89-92 builds them
    anchors = np.array([[116, 90], [156, 198], [373, 326], [30, 61], [62, 45], [59, 119], [10, 13], [16, 30], [33, 23]], np.float32)
    gtwh = np.exp(rng.uniform(np.log(2), np.log(416), size=(200, 2))).astype(np.float32)
    sa = np.concatenate([-anchors / 2, anchors / 2], 1).astype(np.float32)
    sg = np.concatenate([-gtwh / 2, gtwh / 2], 1).astype(np.float32)
    iou_ag = mod.bbox_iou(sa.astype(np.float64), sg.astype(np.float64))
    np.savez_compressed(OUT, a=a, b=b, iou=iou, iou32=iou32, shift_anchor=sa, shift_gt=sg, iou_anchor_gt=iou_ag)
    print("wrote", OUT, iou.shape, iou_ag.shape)


if __name__ == "__main__":
    sys.exit(main())
