"""Independent cross-checks of the box_nms / box_iou restatement (oracle/ASSUMPTIONS.md): torchvision's CPU operators implement
the same published rule (suppress iff inter / (area_a + area_b - inter) > thresh, greedy in score order) and were not written
by us.  Tie-free inputs (distinct scores), corner format, class-aware = torchvision.ops.nms per class."""
import numpy as np
import pytest
import torch

from oracle import ref_nms
from tests.util import random_dets

tv = pytest.importorskip("torchvision")
from torchvision.ops import box_iou as tv_box_iou, nms as tv_nms   # noqa: E402


def _tie_free(rng, B, N, num_class):
    d = random_dets(rng, B, N, num_class=num_class, tie_frac=0.0)
    for b in range(B):                                   # distinct scores: the order is unambiguous for any sort
        d[b, :, 1] = rng.permutation(N).astype(np.float32) / N + 0.02
    return d


def _tv_box_nms(d, thresh, valid, topk, class_aware):
    """box_nms semantics assembled from torchvision pieces: valid filter, descending order, top-k cut, per-class greedy NMS,
    survivors in score order."""
    recs = []
    for img in d:
        idx = np.nonzero(img[:, 1] > valid)[0]
        order = idx[np.argsort(-img[idx, 1], kind="stable")]
        if topk > 0:
            order = order[:topk]
        boxes = torch.from_numpy(img[order, 2:6])
        scores = torch.from_numpy(img[order, 1])
        keep = []
        if class_aware:
            cls = img[order, 0].astype(np.int64)
            for c in np.unique(cls):
                m = np.nonzero(cls == c)[0]
                k = tv_nms(boxes[m], scores[m], thresh).numpy()
                keep.extend(m[k].tolist())
        else:
            keep = tv_nms(boxes, scores, thresh).numpy().tolist()
        keep = sorted(keep)                              # positions in `order` = rank order
        rec = np.full(img.shape[0], -1, np.int32)
        rec[:len(keep)] = order[keep]
        recs.append(rec)
    return np.stack(recs)


@pytest.mark.parametrize("seed,N,topk,C", [(0, 300, 100, 3), (1, 2000, 400, 20), (2, 5000, 400, 7), (3, 800, -1, 1)])
@pytest.mark.parametrize("force", [False, True])
def test_box_nms_matches_torchvision(seed, N, topk, C, force):
    rng = np.random.RandomState(seed)
    d = _tie_free(rng, 2, N, C)
    _, rec = ref_nms.box_nms(d, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0, force_suppress=force,
                             return_record=True)
    exp = _tv_box_nms(d, 0.45, 0.01, topk, class_aware=not force)
    np.testing.assert_array_equal(rec, exp)


def test_box_iou_matches_torchvision():
    rng = np.random.RandomState(4)
    a = random_dets(rng, 1, 64)[0, :, 2:]
    b = random_dets(rng, 1, 48)[0, :, 2:]
    got = ref_nms.box_iou(a, b)
    exp = tv_box_iou(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    np.testing.assert_allclose(got, exp, rtol=1e-6, atol=1e-7)
