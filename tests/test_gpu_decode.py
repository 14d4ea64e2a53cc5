"""YOLOOutputV3 decode on the GPU vs the oracle, fp32: 1e-5 relative (scores) / 1e-5 of the image
size (box corners, which are differences of O(size) terms)."""
import numpy as np
import pytest
import torch

from oracle import ref_head

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def box_close(a, b, size):
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=RTOL * size)


@pytest.mark.parametrize("C,H,W,stride,anchors", [(20, 13, 13, 32, [116, 90, 156, 198, 373, 326]),
                                                  (80, 19, 19, 32, [116, 90, 156, 198, 373, 326]),
                                                  (3, 7, 11, 8, [10, 13, 16, 30, 33, 23]),
                                                  (285, 4, 5, 16, [30, 61, 62, 45, 59, 119])])
def test_decode_modes(C, H, W, stride, anchors):
    import viddet_b200
    rng = np.random.RandomState(C)
    pred = (rng.standard_normal((2, 3 * (5 + C), H, W)) * 1.5).astype(np.float32)
    blk = viddet_b200.YOLOOutputV3(0, C, anchors, stride)
    pt = torch.from_numpy(pred).cuda()
    det = blk.decode(pt).cpu().numpy()
    ref = ref_head.decode(pred, anchors, stride, C)
    assert det.shape == ref.shape
    np.testing.assert_array_equal(det[..., 0], ref[..., 0])
    np.testing.assert_allclose(det[..., 1], ref[..., 1], rtol=RTOL, atol=1e-9)
    box_close(det[..., 2:], ref[..., 2:], stride * max(H, W))
    outs = blk.decode(pt, training=True)
    refs = ref_head.decode(pred, anchors, stride, C, mode="train")
    box_close(outs[0].cpu().numpy(), refs[0], stride * max(H, W))
    for o, r in zip(outs[1:5], refs[1:5]):
        np.testing.assert_array_equal(o.cpu().numpy(), r)        # raw logits are pure copies
    np.testing.assert_array_equal(outs[5].cpu().numpy(), refs[5])
    np.testing.assert_array_equal(outs[6].cpu().numpy(), refs[6])
    ag = viddet_b200.YOLOOutputV3(0, C, anchors, stride, agnostic=True).decode(pt).cpu().numpy()
    ref_ag = ref_head.decode(pred, anchors, stride, C, mode="agnostic")
    np.testing.assert_allclose(ag[..., :2], ref_ag[..., :2], rtol=RTOL)
    box_close(ag[..., 2:], ref_ag[..., 2:], stride * max(H, W))


def test_decode_zero_logit_kat_and_nan_ids():
    import viddet_b200
    C, H, W = 4, 3, 5
    pred = np.zeros((1, 3 * (5 + C), H, W), np.float32)
    pred[0, 5, 0, 0] = np.nan
    blk = viddet_b200.YOLOOutputV3(0, C, [30, 61, 62, 45, 59, 119], 16)
    det = blk.decode(torch.from_numpy(pred).cuda()).cpu().numpy()
    assert np.isnan(det[0, 0, 0]) and np.isnan(det[0, 0, 1])     # ids = score*0 + c (yolo3.py:194)
    np.testing.assert_allclose(det[0, 1:, 1], 0.25, rtol=1e-6)
    np.testing.assert_allclose(det[0, 1, 2:], [8 - 31, 8 - 22.5, 8 + 31, 8 + 22.5], rtol=1e-6)


def test_repack_layout():
    import viddet_b200
    x = torch.randn(3, 70, 5, 9, device="cuda")
    y = viddet_b200.to_nhwc_bf16(x)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y.float(), x.to(torch.bfloat16).float(), rtol=0, atol=0)


def test_postprocess_detections_bit_exact():
    """detect()'s host loop on device (8f row 4): clip, id >= 0 compaction in order, /S, id truncation."""
    import viddet_b200
    from oracle import ref_post
    rng = np.random.RandomState(2)
    B, T, post, S = 3, 5, 100, 416
    ids = rng.randint(-1, 20, size=(B, T, post, 1)).astype(np.float32)
    ids[0, 0] = -1                                            # an image with no detection
    ids[1, 2] = 7                                             # and a full one
    scores = rng.uniform(0, 1, size=(B, T, post, 1)).astype(np.float32)
    boxes = rng.uniform(-60, S + 60, size=(B, T, post, 4)).astype(np.float32)
    ref_rows, ref_cnt = ref_post.postprocess(ids, scores, boxes, S)
    rows, cnt = viddet_b200.postprocess_detections(torch.from_numpy(ids).cuda(), torch.from_numpy(scores).cuda(),
                                                   torch.from_numpy(boxes).cuda(), S)
    np.testing.assert_array_equal(cnt.cpu().numpy(), ref_cnt)
    np.testing.assert_array_equal(rows.cpu().numpy(), ref_rows)
    assert ref_cnt[0] == 0 and ref_cnt[T + 2] == post


def test_feature_stream_from_npy_files(tmp_path):
    """8f row 3: `<id>_F{1,2,3}.npy` (fp32 NCHW, extract_base_features.py:153-155) -> channels-last bf16 tips in the head's
    scale order, bit-identical to converting the loaded arrays directly; short last batch; windows."""
    import viddet_b200
    rng = np.random.RandomState(4)
    ids = ["clip/%06d" % i for i in range(7)]
    (tmp_path / "clip").mkdir()
    shapes = {"_F1.npy": (256, 8, 8), "_F2.npy": (512, 4, 4), "_F3.npy": (1024, 2, 2)}
    data = {}
    for fid in ids:
        for suf, shp in shapes.items():
            a = rng.standard_normal(shp).astype(np.float32)
            np.save(str(tmp_path / (fid + suf)), a)
            data[(fid, suf)] = a
    seen = []
    for got_ids, tips in viddet_b200.FeatureStream(str(tmp_path), ids, batch=3):
        assert [t.shape[1] for t in tips] == [1024, 512, 256]
        for k, suf in enumerate(("_F3.npy", "_F2.npy", "_F1.npy")):
            ref = viddet_b200.to_nhwc_bf16(torch.from_numpy(np.stack([data[(f, suf)] for f in got_ids])).cuda())
            assert tips[k].is_contiguous(memory_format=torch.channels_last)
            assert torch.equal(tips[k].view(torch.int16), ref.view(torch.int16))
        seen += got_ids
    assert seen == ids
    win = list(viddet_b200.FeatureStream(str(tmp_path), ids[:6], batch=1, window=3))
    assert len(win) == 2 and tuple(win[0][1][0].shape) == (1, 3, 1024, 2, 2)


# ------------------------------------------------------------------ hierarchical_nms (detect_yolo3.py:736-789)
def _hier_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hier_nms_golden.npz"))
    for ci in range(int(g["n_cases"])):
        pre = "c%d_" % ci
        yield ci, {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}


def test_hierarchical_nms_matches_reference_golden_bit_exact():
    """Golden vectors = outputs of the reference's own hierarchical_nms executed on these inputs
    (tests/golden/make_golden_hier_nms.py); bit-exact rows and counts."""
    import viddet_b200
    for ci, c in _hier_cases():
        ov, conf, lvl = c["params"]
        tree = viddet_b200.ClassTree(c["levels"], c["parent"], c["branch"])
        out, cnt = viddet_b200.hierarchical_nms(torch.from_numpy(c["rows"]).cuda(), torch.from_numpy(c["counts"]).cuda(), tree,
                                                ov_thresh=float(ov), conf_thresh=float(conf), level_thresh=int(lvl))
        np.testing.assert_array_equal(cnt.cpu().numpy(), c["out_counts"], err_msg="case %d" % ci)
        np.testing.assert_array_equal(out.cpu().numpy(), c["out_rows"], err_msg="case %d" % ci)


def test_hierarchical_nms_vs_oracle_random_and_tree_builder():
    import viddet_b200
    from oracle import ref_post
    rng = np.random.RandomState(77)
    C = 60
    names = ["n%03d" % i for i in range(C)]
    parents = {nm: ("ROOT" if i < 4 else names[rng.randint(0, i)]) for i, nm in enumerate(names)}
    tree = viddet_b200.ClassTree.from_parents(names, parents)
    F, post = 37, 100
    rows = np.full((F, post, 6), -1.0, np.float32); counts = rng.randint(0, post + 1, size=F).astype(np.int32)
    for f in range(F):
        n = counts[f]
        ctr = rng.uniform(50, 350, size=(max(1, n // 5), 2))[rng.randint(0, max(1, n // 5), size=n)] + rng.normal(0, 3, size=(n, 2))
        wh = rng.uniform(30, 120, size=(n, 2))
        rows[f, :n] = np.concatenate([rng.randint(0, C, size=(n, 1)), rng.uniform(0, 1, size=(n, 1)), ctr - wh / 2, ctr + wh / 2], 1)
    lv, pa, br = tree.levels.cpu().numpy(), tree.parent.cpu().numpy(), tree.branch.cpu().numpy()
    for lvl, ov, conf in [(10, 0.5, 0.0), (2, 0.4, 0.3), (1, 0.6, 0.0)]:
        ref, rcnt = ref_post.hierarchical_nms(rows, counts, lv, pa, br, ov, conf, lvl)
        out, cnt = viddet_b200.hierarchical_nms(torch.from_numpy(rows).cuda(), torch.from_numpy(counts).cuda(), tree, ov, conf, lvl)
        np.testing.assert_array_equal(cnt.cpu().numpy(), rcnt)
        np.testing.assert_array_equal(out.cpu().numpy(), ref)
    with pytest.raises(ValueError):
        viddet_b200.hierarchical_nms(torch.from_numpy(rows).cuda(), torch.from_numpy(counts).cuda(), tree, level_thresh=0)
    # chained after the device post-processing of detect()
    ids = torch.from_numpy(rows[..., 0:1]).cuda(); sc = torch.from_numpy(rows[..., 1:2]).cuda(); bb = torch.from_numpy(rows[..., 2:]).cuda()
    prow, pcnt = viddet_b200.postprocess_detections(ids, sc, bb, size=416)
    out, cnt = viddet_b200.hierarchical_nms(prow, pcnt, tree)
    ref, rcnt = ref_post.hierarchical_nms(prow.cpu().numpy(), pcnt.cpu().numpy(), lv, pa, br)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
