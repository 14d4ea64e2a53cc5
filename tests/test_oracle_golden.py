"""Oracle vs golden vectors produced by the reference's own importable helper
(utils/bbox.py::bbox_iou, see tests/golden/make_golden_bbox_iou.py)."""
import os

import numpy as np

from oracle import ref_nms

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "bbox_iou_golden.npz"))


def test_box_iou_matches_reference_helper():
    iou = ref_nms.box_iou(G["a"], G["b"])
    np.testing.assert_allclose(iou, G["iou"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(np.diag(iou[:8, :8]), 1.0, rtol=1e-6)
    assert (iou[8:12, 8:12].diagonal() == 0).all()


def test_anchor_match_matches_reference_helper():
    iou = ref_nms.box_iou(G["shift_anchor"], G["shift_gt"])
    np.testing.assert_allclose(iou, G["iou_anchor_gt"], rtol=2e-5, atol=1e-7)
    # argmax decisions agree wherever the float64 ground truth has a clear winner
    gold = G["iou_anchor_gt"]
    srt = np.sort(gold, axis=0)
    clear = (srt[-1] - srt[-2]) > 1e-5
    np.testing.assert_array_equal(iou.argmax(0)[clear], gold.argmax(0)[clear])
