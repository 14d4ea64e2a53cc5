"""The oracle against golden vectors produced by EXECUTING the reference's own classes (YOLOOutputV3, the prefetch target
generator, the dynamic generator + merger) over a numpy stand-in for the MXNet operators (tests/golden/make_golden_ref_exec.py,
tests/golden/mx_shim.py).  Pins the reference's own logic -- slicing, row order, the per-GT loop, index math, _slice, merges."""
import os

import numpy as np

from oracle import ref_head, ref_loss, ref_targets

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_exec_golden.npz"))
ANCHORS = [[116, 90, 156, 198, 373, 326], [30, 61, 62, 45, 59, 119], [10, 13, 16, 30, 33, 23]]
STRIDES = [32, 16, 8]


def test_decode_matches_executed_reference():
    for ci in range(int(G["n_dec"])):
        pre = "dec%d_" % ci
        C, si, H, W, Cin, B = [int(v) for v in G[pre + "meta"]]
        x, w, b = G[pre + "x"], G[pre + "w"], G[pre + "b"]
        det = ref_head.yolo_output_v3(x, w, b, ANCHORS[si], STRIDES[si], C, "infer")
        gold = G[pre + "infer"]
        assert det.shape == gold.shape == (B, C * H * W * 3, 6)
        np.testing.assert_array_equal(det[..., 0], gold[..., 0])                       # class ids: the row order itself
        np.testing.assert_allclose(det[..., 1], gold[..., 1], rtol=1e-5, atol=1e-7)    # conv summation order differs
        np.testing.assert_allclose(det[..., 2:], gold[..., 2:], rtol=1e-5, atol=2e-4)
        ag = ref_head.yolo_output_v3(x, w, b, ANCHORS[si], STRIDES[si], C, "agnostic")
        np.testing.assert_allclose(ag, G[pre + "agnostic"], rtol=1e-5, atol=2e-4)
        tr = ref_head.yolo_output_v3(x, w, b, ANCHORS[si], STRIDES[si], C, "train")
        for got, k in zip(tr, ("bbox", "raw_centers", "raw_scales", "objness", "class_pred", "anchors", "offsets")):
            gold = G[pre + "train_" + k]
            assert np.shape(got) == gold.shape, k
            np.testing.assert_allclose(got, gold, rtol=1e-5, atol=2e-4, err_msg=k)


def _tg_inputs(pre):
    C, B, M, size, multi, mix = [int(v) for v in G[pre + "meta"]]
    img_shape, xs_shapes, anchors, offsets = ref_targets.default_generator_inputs(size)
    return C, (B,) + tuple(img_shape[1:]), xs_shapes, anchors, offsets, G[pre + "gt"], G[pre + "ids"], (G[pre + "mix"] if mix else None)


def test_prefetch_targets_match_executed_reference():
    names = ("objectness", "center", "scale", "weights", "class")
    for ci in range(int(G["n_tg"])):
        pre = "tg%d_" % ci
        C, img_shape, xs_shapes, anchors, offsets, gt, ids, mix = _tg_inputs(pre)
        # same NumPy scalar semantics as the executed reference: bit-exact, every tensor
        res = ref_targets.prefetch_targets(img_shape, xs_shapes, anchors, offsets, gt, ids, mix, num_class=C, promotion="nep50")
        for k, r in zip(names, res):
            np.testing.assert_array_equal(r, G[pre + k], err_msg="%s %s" % (pre, k))
        # the era's promotion (what the product implements): same assignments and class rows, values within ~1 ulp
        leg = ref_targets.prefetch_targets(img_shape, xs_shapes, anchors, offsets, gt, ids, mix, num_class=C)
        np.testing.assert_array_equal(leg[0], G[pre + "objectness"])
        np.testing.assert_array_equal(leg[4], G[pre + "class"])
        for i in (1, 2, 3):
            np.testing.assert_array_equal(leg[i] != 0, G[pre + names[i]] != 0)
            np.testing.assert_allclose(leg[i], G[pre + names[i]], rtol=5e-7, atol=4e-6)
        assert (G[pre + "objectness"] > 0).sum() > 0


def test_target_merger_matches_executed_reference():
    names = ("objectness", "center", "scale", "weights", "class", "class_mask")
    for ci in range(int(G["n_tg"])):
        pre = "tg%d_" % ci
        C = int(G[pre + "meta"][0])
        pf = [G[pre + k] for k in names[:5]]
        res = ref_loss.target_merge(G[pre + "preds"], G[pre + "gt"], *pf, num_class=C, ignore_iou_thresh=0.7)
        for k, r in zip(names, res):
            np.testing.assert_array_equal(np.asarray(r, np.float32), G[pre + "merged_" + k], err_msg="%s %s" % (pre, k))
        assert (G[pre + "merged_objectness"] < 0).sum() > 0          # some predictions are ignored (IoU > 0.7, not matched)


def test_time_distributed_and_pooling_match_executed_reference():
    from oracle import ref_temporal
    C, si, H, W, Cin, B, T = [int(v) for v in G["td_meta"]]
    x = G["td_x"]
    fn = lambda z: ref_head.yolo_output_v3(z, G["td_w"], G["td_b"], ANCHORS[si], STRIDES[si], C, "infer")
    out = ref_temporal.time_distributed(fn, x)
    assert out.shape == G["td_out"].shape == (B, T, C * H * W * 3, 6)
    np.testing.assert_array_equal(out[..., 0], G["td_out"][..., 0])
    np.testing.assert_allclose(out, G["td_out"], rtol=1e-5, atol=2e-4)
    np.testing.assert_array_equal(ref_temporal.temporal_pooling(x, "max"), G["pool_max"])
    np.testing.assert_allclose(ref_temporal.temporal_pooling(x, "mean"), G["pool_mean"], rtol=1e-6, atol=1e-7)


def neck_params():
    """Layer parameters of the executed YOLOV3 in execution order -> (blocks, transitions, prediction convs)."""
    from tests.util import replay_shim_params
    ps = replay_shim_params(77, G["neck_param_kinds"], G["neck_param_shapes"], G["neck_param_bias"])
    it = iter(ps)

    def cell():
        c, b = next(it), next(it)
        assert c[0] == "conv" and b[0] == "bn" and c[2] is None
        return dict(weight=c[1], gamma=b[1], beta=b[2], mean=b[3], var=b[4])
    blocks, transitions, preds = [], [], []
    for i in range(3):                       # yolo3.py:496-519: block i, output i, then transition i
        blocks.append([cell() for _ in range(6)])
        p = next(it)
        assert p[0] == "conv" and p[2] is not None
        preds.append((p[1], p[2]))
        if i < 2:
            transitions.append(cell())
    assert next(it, None) is None
    return blocks, transitions, preds


def test_yolov3_forward_after_stages_matches_executed_reference():
    """YOLOV3.hybrid_forward (inference) executed from the reference source: detection blocks, transitions, x2 upsample +
    concat order, reversed routes, per-scale outputs, concat, box_nms parameters and the post_nms slice."""
    from oracle import ref_block, ref_nms
    C, B, _ = [int(v) for v in G["neck_meta"]]
    blocks, transitions, preds = neck_params()
    feats = [G["neck_feat%d" % i] for i in range(3)]
    tips = ref_block.yolo3_neck_tips(feats, blocks, transitions)
    assert [t.shape[1] for t in tips] == [256, 256, 256] and [t.shape[-1] for t in tips] == [4, 8, 16]
    det = ref_head.head_detections(tips, [p[0] for p in preds], [p[1] for p in preds], C)
    gold = G["neck_det"]
    assert det.shape == gold.shape
    np.testing.assert_array_equal(det[..., 0], gold[..., 0])
    np.testing.assert_allclose(det[..., 1], gold[..., 1], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(det[..., 2:], gold[..., 2:], rtol=2e-4, atol=2e-3)
    out = ref_nms.box_nms(gold, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1, coord_start=2)[:, :100]
    np.testing.assert_array_equal(out[..., 0:1], G["neck_ids"])
    np.testing.assert_array_equal(out[..., 1:2], G["neck_scores"])
    np.testing.assert_array_equal(out[..., 2:], G["neck_bboxes"])


def tneck_params():
    from tests.util import replay_shim_params
    ps = replay_shim_params(78, G["tneck_param_kinds"], G["tneck_param_shapes"], G["tneck_param_bias"])
    chk = np.array([float(np.asarray(p[1], np.float64).sum()) for p in ps])
    np.testing.assert_array_equal(chk, G["tneck_param_check"])
    it = iter(ps)

    def cell():
        c, b = next(it), next(it)
        assert c[0] == "conv" and b[0] == "bn" and c[2] is None
        return dict(weight=c[1], gamma=b[1], beta=b[2], mean=b[3], var=b[4])
    blocks, transitions, preds = [], [], []
    for i in range(3):                       # yolo3_temporal.py:448-497: (2+1)D block i (9 cells), output i, then transition i
        blocks.append([cell() for _ in range(9)])
        p = next(it)
        assert p[0] == "conv" and p[2] is not None
        preds.append((p[1], p[2]))
        if i < 2:
            transitions.append(cell())
    assert next(it, None) is None
    return blocks, transitions, preds


def test_yolov3_temporal_forward_matches_executed_reference():
    """YOLOV3Temporal.hybrid_forward (t=5, t_out, conv type 21) executed from the reference source: TimeDistributed stages / output
    layers / transitions, (2+1)D detection blocks (swapaxes around the 3-D convs), 5-D upsample + concat, box_nms over (B,T,rows,6)."""
    from oracle import ref_block, ref_nms, ref_temporal
    C, B, T, _ = [int(v) for v in G["tneck_meta"]]
    blocks, transitions, preds = tneck_params()
    feats = [G["tneck_feat%d" % i] for i in range(3)]
    tips = ref_block.yolo3_neck_tips(feats, blocks, transitions, conv_type="21")
    assert [t.shape for t in tips] == [(B, T, 256, 2, 2), (B, T, 256, 4, 4), (B, T, 256, 8, 8)]
    head = lambda *frames: ref_head.head_detections(list(frames), [p[0] for p in preds], [p[1] for p in preds], C)
    det = head(*[t.reshape((B * T,) + t.shape[2:]) for t in tips]).reshape(B, T, -1, 6)        # TimeDistributed(output), concat dim -2
    gold = G["tneck_det"]
    assert det.shape == gold.shape
    np.testing.assert_array_equal(det[..., 0], gold[..., 0])
    np.testing.assert_allclose(det[..., 1], gold[..., 1], rtol=3e-4, atol=1e-6)
    np.testing.assert_allclose(det[..., 2:], gold[..., 2:], rtol=3e-4, atol=3e-3)
    out = ref_nms.box_nms(gold, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1, coord_start=2)[..., :100, :]
    np.testing.assert_array_equal(out[..., 0:1], G["tneck_ids"])
    np.testing.assert_array_equal(out[..., 1:2], G["tneck_scores"])
    np.testing.assert_array_equal(out[..., 2:], G["tneck_bboxes"])


def train_outputs_oracle(tips, preds, C):
    """yolo3.py:498-509,532-535 on oracle pieces: per-scale train-mode 7-tuples, reshape((0,-3,-1)), concat over the scales."""
    outs = [ref_head.yolo_output_v3(t, w, b, ANCHORS[i], STRIDES[i], C, "train") for i, (t, (w, b)) in enumerate(zip(tips, preds))]
    B = tips[0].shape[0]
    cat = lambda j, last: np.concatenate([np.asarray(o[j]).reshape(B, -1, last) for o in outs], axis=1)
    return cat(0, 4), cat(1, 2), cat(2, 2), cat(3, 1), cat(4, C)


def test_training_branch_raw_predictions_match_executed_reference():
    from oracle import ref_block
    C, B, _ = [int(v) for v in G["neck_meta"]]
    blocks, transitions, preds = neck_params()
    feats = [G["neck_feat%d" % i] for i in range(3)]
    tips = ref_block.yolo3_neck_tips(feats, blocks, transitions)
    got = train_outputs_oracle(tips, preds, C)
    for g, k in zip(got, ("box_preds", "box_centers", "box_scales", "objness", "class_pred")):
        gold = G["neck_train_" + k]
        assert g.shape == gold.shape, k
        np.testing.assert_allclose(g, gold, rtol=3e-4, atol=3e-3, err_msg=k)
    np.testing.assert_array_equal(G["neck_train_fmap_shapes"], [[1, 1, 4, 4], [1, 1, 8, 8], [1, 1, 16, 16]])
