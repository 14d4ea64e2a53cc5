"""Edge cases of the hot path on the GPU: empty and one-element inputs, everything filtered, the reference's own batch-1 case
(BASELINE configs[0]), maximum top-k, ragged maps -- each against the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_head, ref_nms, ref_post, ref_targets
from tests.util import bf16_round, make_pred_weights, make_tips, random_dets

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_head(C, ws, bs, **kw):
    import viddet_b200
    head = viddet_b200.YOLOV3Head(C, **kw)
    for o, w, b in zip(head.yolo_outputs, ws, bs):
        o.prediction.set_data(torch.from_numpy(w), torch.from_numpy(b))
    return head


def test_box_nms_degenerate_shapes():
    import viddet_b200
    kw = dict(overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1, coord_start=2)
    rng = np.random.RandomState(0)
    # one row per image; a single image; more top-k than rows
    for shape in [(3, 1), (1, 17), (2, 5)]:
        d = random_dets(rng, shape[0], shape[1], num_class=2)
        out, rec = viddet_b200.box_nms(cuda(d), return_record=True, **kw)
        ref, rrec = ref_nms.box_nms(d, return_record=True, **kw)
        np.testing.assert_array_equal(out.cpu().numpy(), ref)
        np.testing.assert_array_equal(rec.cpu().numpy(), rrec)
    # everything at or below the valid threshold (strict >): all rows -1
    d = random_dets(rng, 2, 50, num_class=3)
    d[..., 1] = 0.01
    out = viddet_b200.box_nms(cuda(d), **kw).cpu().numpy()
    assert (out == -1).all()
    np.testing.assert_array_equal(out, ref_nms.box_nms(d, **kw))
    # zero images / zero rows: same (empty) shape back, no kernel work
    for shape in [(0, 40, 6), (3, 0, 6)]:
        e = torch.empty(shape, device="cuda")
        assert tuple(viddet_b200.box_nms(e, **kw).shape) == shape


def test_head_batch_one_is_the_reference_cpu_case():
    """BASELINE configs[0]: one synthetic 416 x 416 frame, batch 1 (what detect_yolo3.py runs)."""
    rng = np.random.RandomState(10)
    C = 20
    tips = make_tips(rng, 1)
    ws, bs = make_pred_weights(rng, C)
    head = build_head(C, ws, bs)
    tt = [cuda(t) for t in tips]
    for _ in range(3):                                   # first call = exact path on a fresh workspace, then speculative steady state
        ids, scores, boxes, keep = [t.cpu().numpy() for t in head(tt, return_keep=True)]
        det = head.detections(tt)
        import viddet_b200
        out, rec = viddet_b200.box_nms(det, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, return_record=True)
        np.testing.assert_array_equal(keep, rec[:, :100].cpu().numpy())
    odet = ref_head.head_detections(tips, ws, bs, C)
    oout, orec = ref_nms.box_nms(odet, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1, coord_start=2,
                                 return_record=True)
    same = keep == orec[:, :100]
    assert same.mean() >= 0.97
    np.testing.assert_allclose(scores[..., 0][same], oout[:, :100, 1][same], rtol=1e-3)


def test_head_everything_filtered_and_max_topk():
    import viddet_b200
    rng = np.random.RandomState(11)
    C, B = 20, 2
    tips = make_tips(rng, B, size=160)
    ws, bs = make_pred_weights(rng, C)
    for b in bs:
        b.reshape(3, 5 + C)[:, 4] = -30.0                # objectness ~1e-13: every score is below valid_thresh
    head = build_head(C, ws, bs)
    tt = [cuda(t) for t in tips]
    for _ in range(2):
        ids, scores, boxes = head(tt)
        assert (ids == -1).all() and (scores == -1).all() and (boxes == -1).all()
    # the largest top-k the fused path supports (VD_MAX_TOPK), post_nms = top-k
    ws, bs = make_pred_weights(rng, C)
    head = build_head(C, ws, bs)
    head.set_nms(0.45, viddet_b200._lib.VD_MAX_TOPK, viddet_b200._lib.VD_MAX_TOPK)
    for _ in range(2):
        ids, scores, boxes, keep = head(tt, return_keep=True)
        out, rec = viddet_b200.box_nms(head.detections(tt), overlap_thresh=0.45, valid_thresh=0.01, topk=viddet_b200._lib.VD_MAX_TOPK,
                                       id_index=0, return_record=True)
        assert torch.equal(keep, rec[:, :viddet_b200._lib.VD_MAX_TOPK])
    head.set_nms(0.45, viddet_b200._lib.VD_MAX_TOPK + 1, 100)
    with pytest.raises(viddet_b200.VidDetError):
        head(tt)


def test_empty_batches_everywhere():
    import viddet_b200
    cell = viddet_b200.ConvBNLReLU(64, 128, 3).initialize()
    y = cell(torch.empty((0, 64, 5, 5), device="cuda"))
    assert tuple(y.shape) == (0, 128, 5, 5)
    up = viddet_b200.upsample_concat(torch.empty((0, 64, 2, 2), device="cuda"), torch.empty((0, 64, 4, 4), device="cuda"))
    assert tuple(up.shape) == (0, 128, 4, 4)
    rows, counts = viddet_b200.postprocess_detections(torch.empty((0, 100, 1), device="cuda"), torch.empty((0, 100, 1), device="cuda"),
                                                      torch.empty((0, 100, 4), device="cuda"), size=416)
    assert tuple(rows.shape) == (0, 100, 6) and counts.numel() == 0
    tree = viddet_b200.ClassTree([1, 2], [-1, 0], [[1, 1], [1, 1]])
    o, c = viddet_b200.hierarchical_nms(rows, counts, tree)
    assert tuple(o.shape) == (0, 100, 6)
    head = viddet_b200.YOLOV3Head(20).initialize()
    ids, scores, boxes = head([torch.empty((0, c, 416 // st, 416 // st), device="cuda") for c, st in zip((1024, 512, 256), (32, 16, 8))])
    assert tuple(ids.shape) == (0, 100, 1) and tuple(boxes.shape) == (0, 100, 4)
    # images without a single detection
    rows = torch.full((3, 10, 6), -1.0, device="cuda"); counts = torch.zeros(3, dtype=torch.int32, device="cuda")
    o, c = viddet_b200.hierarchical_nms(rows, counts, tree)
    assert (c == 0).all() and (o == -1).all()


def test_targets_without_ground_truth_and_single_box():
    import viddet_b200
    C, size = 20, 416
    img, xs, anchors, offsets = ref_targets.default_generator_inputs(size)
    gen = viddet_b200.YOLOV3PrefetchTargetGenerator(C)
    anc = [torch.from_numpy(a) for a in anchors]
    gt = np.full((2, 5, 4), -1.0, np.float32); ids = np.full((2, 5, 1), -1.0, np.float32)
    outs = [o.cpu().numpy() for o in gen(img, xs, anc, offsets, cuda(gt), cuda(ids))]
    ref = ref_targets.prefetch_targets(img, xs, anchors, offsets, gt, ids, None, num_class=C)
    for o, r in zip(outs, ref):
        np.testing.assert_array_equal(o, r)
    assert (outs[0] == 0).all() and (outs[4] == -1).all()
    gt[1, 0] = [100.3, 50.2, 180.9, 220.4]; ids[1, 0, 0] = 7          # one box in the second image only
    outs = [o.cpu().numpy() for o in gen(img, xs, anc, offsets, cuda(gt), cuda(ids))]
    ref = ref_targets.prefetch_targets(img, xs, anchors, offsets, gt, ids, None, num_class=C)
    np.testing.assert_array_equal(outs[0], ref[0]); np.testing.assert_array_equal(outs[4], ref[4])
    assert outs[0][0].sum() == 0 and outs[0][1].sum() == 1


def test_postprocess_single_row_and_all_invalid():
    import viddet_b200
    ids = np.array([[[3.0]], [[-1.0]]], np.float32); sc = np.array([[[0.5]], [[-1.0]]], np.float32)
    bb = np.array([[[-4.0, 10.0, 500.0, 300.0]], [[-1.0, -1.0, -1.0, -1.0]]], np.float32)
    rows, counts = viddet_b200.postprocess_detections(cuda(ids), cuda(sc), cuda(bb), size=416)
    r, c = ref_post.postprocess(ids, sc, bb, 416)
    np.testing.assert_array_equal(rows.cpu().numpy(), r); np.testing.assert_array_equal(counts.cpu().numpy(), c)
    assert counts.tolist() == [1, 0]
