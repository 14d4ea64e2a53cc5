"""fp32-parity mode of the fused head (VD_PREC_FP32_SPLIT, include/viddet_b200.h): the reference head is fp32 end to end
(yolo3.py:62,157-199); with hi/lo bf16 operand planes and three products per term the tcgen05 path must land within 1e-5
relative of the fp32 oracle on arbitrary fp32 inputs (NOT pre-rounded to bf16), and reproduce its NMS keep rows.

Tolerances (north star: "1e-5 (fp32)"): scores rtol 1e-5; box corners 1e-5 of the box's own extent max(|corner|, image size)
(x1 = cx - w/2 cancels, so a corner near 0 has no meaningful relative error of its own; w = exp(tw) * anchor reaches thousands of
pixels under random weights); raw conv outputs 1e-5 of the tensor's max magnitude.  'bf16x2' (two planes): 1e-4."""
import numpy as np
import pytest
import torch

from oracle import ref_head, ref_nms
from tests.util import ANCHORS, keep_agreement, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-5
TOL = {"fp32": 1e-5, "bf16x2": 1e-4}


def box_err(got, ref, size):
    """max |corner error| / max(|corners of that box|, image size)."""
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    ext = np.maximum(np.abs(ref).max(axis=-1, keepdims=True), float(size))
    return float((np.abs(got - ref) / ext).max())


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def fp32_inputs(rng, B, C, size, bias_scale=0.1):
    """Full-mantissa fp32 tips (leaky_relu(N(0,1))) and U(-0.07, 0.07) weights: nothing is representable in bf16."""
    tips, ws, bs = [], [], []
    n = 3 * (5 + C)
    for c, s in zip([1024, 512, 256], [32, 16, 8]):
        h = size // s
        x = rng.standard_normal((B, c, h, h)).astype(np.float32)
        tips.append(np.where(x > 0, x, np.float32(0.1) * x).astype(np.float32))
        ws.append(rng.uniform(-0.07, 0.07, (n, c, 1, 1)).astype(np.float32))
        bs.append(rng.uniform(-bias_scale, bias_scale, n).astype(np.float32))
    return tips, ws, bs


def build_head(C, ws, bs, **kw):
    import viddet_b200
    head = viddet_b200.YOLOV3Head(C, **kw)
    for o, w, b in zip(head.yolo_outputs, ws, bs):
        o.prediction.set_data(torch.from_numpy(w), torch.from_numpy(b))
    return head


@pytest.mark.parametrize("C,H,W,Cin", [(20, 13, 13, 1024), (20, 26, 26, 512), (30, 52, 52, 256), (80, 19, 19, 1024), (3, 5, 7, 64)])
def test_pred_conv_fp32_split_vs_float64(C, H, W, Cin):
    import viddet_b200
    rng = np.random.RandomState(C + H)
    x = rng.standard_normal((2, Cin, H, W)).astype(np.float32)
    n = 3 * (5 + C)
    w = rng.uniform(-0.07, 0.07, (n, Cin, 1, 1)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, n).astype(np.float32)
    blk = viddet_b200.YOLOOutputV3(0, C, ANCHORS[0], 32, precision="fp32")
    blk.prediction.set_data(torch.from_numpy(w), torch.from_numpy(b))
    pred = blk.predict(cuda(x)).cpu().numpy()
    ref64 = np.einsum("nk,bkp->bnp", w.reshape(n, Cin).astype(np.float64), x.reshape(2, Cin, H * W).astype(np.float64))
    ref64 = (ref64 + b.astype(np.float64)[None, :, None]).reshape(pred.shape)
    ref32 = ref_head.conv1x1(x, w, b)
    scale = np.abs(ref64).max()
    e_dev, e_f32 = np.abs(pred - ref64).max() / scale, np.abs(ref32 - ref64).max() / scale
    blk2 = viddet_b200.YOLOOutputV3(0, C, ANCHORS[0], 32, precision="bf16x2")
    blk2.prediction.set_data(torch.from_numpy(w), torch.from_numpy(b))
    e_2 = np.abs(blk2.predict(cuda(x)).cpu().numpy() - ref64).max() / scale
    print("pred conv C=%d Cin=%d: err of max |y|: 3 planes %.2e, 2 planes %.2e (numpy fp32 conv itself: %.2e)" % (C, Cin, e_dev, e_2, e_f32))
    assert e_dev <= 2e-6 and e_2 <= 2e-5
    # the bf16 path on the same fp32 inputs is ~100x further away: the mode is doing something
    blk16 = viddet_b200.YOLOOutputV3(0, C, ANCHORS[0], 32)
    blk16.prediction.set_data(torch.from_numpy(w), torch.from_numpy(b))
    e16 = np.abs(blk16.predict(cuda(x)).cpu().numpy() - ref64).max() / scale
    assert e16 > 10 * e_dev


@pytest.mark.parametrize("precision", ["fp32", "bf16x2"])
@pytest.mark.parametrize("C,size,B", [(20, 416, 2), (30, 224, 2), (80, 320, 1)])
def test_detections_fp32_vs_oracle(C, size, B, precision):
    """The materialised (B, rows, 6) tensor of yolo3.py:523 from fp32 NCHW inputs: ids exact, scores 1e-5 relative, boxes
    1e-5 of their extent (1e-4 for the two-plane mode)."""
    rng = np.random.RandomState(C + size)
    tips, ws, bs = fp32_inputs(rng, B, C, size)
    head = build_head(C, ws, bs, precision=precision)
    det = head.detections([cuda(t) for t in tips]).cpu().numpy()
    ref = ref_head.head_detections(tips, ws, bs, C)
    np.testing.assert_array_equal(det[..., 0], ref[..., 0])
    es = rel_err(det[..., 1], ref[..., 1], 1e-3)
    eb = box_err(det[..., 2:], ref[..., 2:], size)
    print("detections %s C=%d size=%d: score rel err %.2e, box err %.2e of the box extent" % (precision, C, size, es, eb))
    assert es <= TOL[precision] and eb <= TOL[precision]


def test_fused_fp32_vs_oracle_voc416_keep_rows():
    """The fused call (conv + decode + top-k + NMS) in fp32 mode against the oracle chain on the same fp32 inputs: scores /
    boxes of the returned rows within 1e-5, keep rows identical except where two candidates' oracle scores are closer than
    the 1e-5 noise band (then they may swap ranks)."""
    rng = np.random.RandomState(0)
    C, B, size = 20, 4, 416
    tips, ws, bs = fp32_inputs(rng, B, C, size)
    head = build_head(C, ws, bs, precision="fp32")
    head.set_nms(0.45, 400, 100)
    tt = [cuda(t) for t in tips]
    for call in range(2):                                    # cold (exact path) and steady state (speculative path)
        ids, scores, boxes, keep = [t.cpu().numpy() for t in head(tt, return_keep=True)]
        det = ref_head.head_detections(tips, ws, bs, C)
        out, rec = ref_nms.box_nms(det, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1,
                                   coord_start=2, return_record=True)
        frac, n_tie = keep_agreement(keep, rec, det, tol=2 * RTOL)
        print("fused fp32 voc416 call %d: %.2f%% identical keep positions, %d near-tie swaps" % (call, 100 * frac, n_tie))
        assert frac >= 0.99
        np.testing.assert_allclose(scores[..., 0], out[:, :100, 1], rtol=2 * RTOL)          # position-wise: a swapped near-tie pair differs by < 2e-5
        same = keep == rec[:, :100]
        np.testing.assert_allclose(scores[..., 0][same], out[:, :100, 1][same], rtol=RTOL)
        assert box_err(boxes[same], out[:, :100, 2:][same], size) <= RTOL
        np.testing.assert_array_equal(ids[..., 0][same], out[:, :100, 0][same])


def test_fused_fp32_tie_free_input_is_identical():
    """A small head whose top-k scores are verifiably separated by more than the noise band (checked on the oracle): the keep
    rows must then be IDENTICAL to the oracle's, position by position."""
    C, B, size, topk = 4, 2, 96, 40
    for seed in range(40):
        rng = np.random.RandomState(100 + seed)
        tips, ws, bs = fp32_inputs(rng, B, C, size, bias_scale=0.5)
        det = ref_head.head_detections(tips, ws, bs, C)
        top = -np.sort(-det[..., 1], axis=1)[:, :topk + 1].astype(np.float64)
        gap = ((top[:, :-1] - top[:, 1:]) / top[:, :-1]).min()
        if gap > 2e-4:
            break
    else:
        pytest.skip("no tie-free seed found")
    head = build_head(C, ws, bs, precision="fp32")
    head.set_nms(0.45, topk, 100)
    ids, scores, boxes, keep = [t.cpu().numpy() for t in head([cuda(t) for t in tips], return_keep=True)]
    out, rec = ref_nms.box_nms(det, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0, score_index=1, coord_start=2,
                               return_record=True)
    np.testing.assert_array_equal(keep, rec[:, :100])
    np.testing.assert_array_equal(ids[..., 0], out[:, :100, 0])
    np.testing.assert_allclose(scores[..., 0], out[:, :100, 1], rtol=RTOL)
    valid = keep >= 0
    assert box_err(boxes[valid], out[:, :100, 2:][valid], size) <= RTOL


def test_fp32_mode_is_bit_identical_between_fused_and_compat_chain():
    """Same guarantee as the bf16 path: head() == box_nms(detections())[:, :100] bit for bit, also in fp32 mode."""
    import viddet_b200
    rng = np.random.RandomState(5)
    C, B, size = 20, 3, 320
    tips, ws, bs = fp32_inputs(rng, B, C, size)
    head = build_head(C, ws, bs, precision="fp32")
    head.set_nms(0.45, 400, 100)
    tt = [cuda(t) for t in tips]
    det = head.detections(tt)
    out, rec = viddet_b200.box_nms(det, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1,
                                   coord_start=2, force_suppress=False, return_record=True)
    for call in range(3):
        ids, scores, boxes, keep = head(tt, return_keep=True)
        assert torch.equal(keep, rec[:, :100])
        assert torch.equal(scores.view(torch.int32), out[:, :100, 1:2].contiguous().view(torch.int32))
        assert torch.equal(boxes.view(torch.int32), out[:, :100, 2:].contiguous().view(torch.int32))


def test_fp32_mode_rejects_unsupported_combinations():
    import viddet_b200
    with pytest.raises(NotImplementedError):
        viddet_b200.YOLOV3Head(20, temporal="cat", k=3, precision="fp32")
    viddet_b200.YOLOV3Head(20, temporal="conv21", precision="fp32")        # the temporal head has a parity mode too
    with pytest.raises(ValueError):
        viddet_b200.YOLOV3Head(20, precision="fp16")


def test_temporal_tip_cell_fp32_vs_oracle():
    """Conv3D((3,1,1)) + BN + LeakyReLU (layers.py:82-89) in the fp32-parity mode: full-mantissa inputs / weights / BN statistics,
    split carriers in and out; the reassembled fp32 output within 1e-5 of the largest activation."""
    import viddet_b200
    from oracle import ref_temporal
    rng = np.random.RandomState(4)
    for (Cc, H, Wd, B, T) in [(256, 13, 13, 2, 5), (512, 6, 5, 1, 5), (128, 20, 20, 2, 3)]:
        x = rng.standard_normal((B, T, Cc, H, Wd)).astype(np.float32)
        w = rng.uniform(-0.07, 0.07, (Cc, Cc, 3, 1, 1)).astype(np.float32)
        gamma = rng.uniform(0.5, 1.5, Cc).astype(np.float32); beta = rng.uniform(-0.2, 0.2, Cc).astype(np.float32)
        mean = rng.uniform(-0.2, 0.2, Cc).astype(np.float32); var = rng.uniform(0.5, 1.5, Cc).astype(np.float32)
        cell = viddet_b200.TemporalTipConv(Cc, precision="fp32")
        cell.set_data(torch.from_numpy(w), gamma, beta, mean, var)
        ys = cell(cuda(x))
        assert isinstance(ys, viddet_b200.SplitF32) and ys.planes == 3
        y = ys.data.float().sum(dim=0).permute(0, 3, 1, 2).reshape(B, T, Cc, H, Wd).cpu().numpy()     # p0 + p1 + p2 (exact in fp32: disjoint mantissa ranges)
        ref = ref_temporal.temporal_conv_bn_lrelu(x, w, gamma, beta, mean, var)
        e = np.abs(y - ref).max() / np.abs(ref).max()
        print("temporal tip cell fp32 mode C=%d: err %.2e of max|y|" % (Cc, e))
        assert e <= 3e-6


def test_temporal_head_conv21_fp32_vs_oracle():
    """cfg 4's head (tip cell at three scales -> pred conv -> decode -> box_nms over (B,T,rows,6)) in the fp32-parity mode, against
    the all-fp32 oracle chain with NO intermediate rounding: scores 1e-5, boxes 1e-5 of their extent, keep rows equal up to near-ties."""
    from oracle import ref_temporal
    rng = np.random.RandomState(6)
    C, B, T, size = 30, 1, 5, 160
    tips5, ws, bs = [], [], []
    n = 3 * (5 + C)
    tw = []
    for c, s in zip([1024, 512, 256], [32, 16, 8]):
        h = size // s
        x = rng.standard_normal((B, T, c, h, h)).astype(np.float32)
        tips5.append(np.where(x > 0, x, np.float32(0.1) * x).astype(np.float32))
        ws.append(rng.uniform(-0.07, 0.07, (n, c, 1, 1)).astype(np.float32))
        bs.append(rng.uniform(-0.1, 0.1, n).astype(np.float32))
        tw.append(rng.uniform(-0.05, 0.05, (c, c, 3, 1, 1)).astype(np.float32))
    head = build_head(C, ws, bs, temporal="conv21", precision="fp32")
    for tc_, w in zip(head.tip_convs, tw):
        tc_.set_data(torch.from_numpy(w))
    head.set_nms(0.45, 400, 100)
    t5 = [cuda(t) for t in tips5]
    det = head.detections(t5).cpu().numpy()
    ones, zeros = (lambda c: np.ones(c, np.float32)), (lambda c: np.zeros(c, np.float32))
    mid = [ref_temporal.temporal_conv_bn_lrelu(t, w, ones(c), zeros(c), zeros(c), ones(c)) for t, w, c in zip(tips5, tw, [1024, 512, 256])]
    ref = ref_head.head_detections([m.reshape((B * T,) + m.shape[2:]) for m in mid], ws, bs, C).reshape(det.shape)
    es = rel_err(det[..., 1], ref[..., 1], 1e-3)
    eb = box_err(det[..., 2:], ref[..., 2:], size)
    print("temporal conv21 head, fp32 mode: score rel err %.2e, box err %.2e of the box extent" % (es, eb))
    assert es <= RTOL and eb <= RTOL
    ids, scores, boxes, keep = head(t5, return_keep=True)
    out, rec = ref_nms.box_nms(ref, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1, coord_start=2,
                               return_record=True)
    frac, n_tie = keep_agreement(keep.reshape(B * T, 100).cpu().numpy(), rec.reshape(B * T, -1), ref.reshape(B * T, -1, 6), tol=2 * RTOL)
    print("temporal conv21 head, fp32 mode: %.2f%% identical keep positions, %d near-tie swaps" % (100 * frac, n_tie))
    assert frac >= 0.99
