"""GPU parity: vd_target_merge / vd_yolo3_loss vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_loss, ref_targets
from tests.util import ANCHORS, make_gt

pytestmark = pytest.mark.gpu


def cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def synth(rng, B, M, C, size=416, multi_hot=False):
    """Prefetched targets from the oracle's target generator + random train-mode predictions around the GTs."""
    gt, ids = make_gt(rng, B, M, size=size, num_class=C, multi_hot=multi_hot, min_count=1)
    hs = [size // s for s in (32, 16, 8)]
    pre = ref_targets.prefetch_targets((B, 3, size, size), [(B, 1, h, h) for h in hs],
                                       [np.asarray(a, np.float32).reshape(1, 1, 3, 2) for a in ANCHORS],
                                       [np.zeros((1, h * h, 1, 2), np.float32) for h in hs], gt, ids, None, num_class=C)
    N = pre[0].shape[1]
    # predicted boxes: random, with a share of them jittered copies of GT boxes so that some IoUs cross the threshold
    xy = rng.uniform(0, size - 40, (B, N, 2)); wh = np.exp(rng.uniform(np.log(8), np.log(200), (B, N, 2)))
    box = np.concatenate([xy, xy + wh], -1).astype(np.float32)
    for b in range(B):
        valid = np.where(gt[b, :, 0] >= 0)[0]
        pick = rng.choice(N, size=N // 20, replace=False)
        box[b, pick] = gt[b, rng.choice(valid, size=len(pick))] + rng.normal(0, 3.0, (len(pick), 4)).astype(np.float32)
    return gt, box, pre


@pytest.mark.parametrize("B,M,C,multi,smooth", [(3, 12, 20, False, False), (2, 100, 30, True, False), (2, 7, 5, False, True)])
def test_target_merge_bit_exact(B, M, C, multi, smooth):
    import viddet_b200
    rng = np.random.RandomState(B * 100 + M)
    gt, box, pre = synth(rng, B, M, C, multi_hot=multi)
    ref = ref_loss.target_merge(box, gt, *pre, C, 0.7, label_smooth=smooth)
    mg = viddet_b200.YOLOV3TargetMerger(C, 0.7)
    mg._label_smooth = smooth
    out = mg(cuda(box), cuda(gt), *[cuda(t) for t in pre])
    assert (ref[0] == -1).any() and (ref[0] == 1).any()
    for o, r in zip(out, ref):
        np.testing.assert_array_equal(o.cpu().numpy(), r)
    dyn = viddet_b200.YOLOV3DynamicTargetGeneratorSimple(C, 0.7)(cuda(box), cuda(gt))
    for o, r in zip(dyn, ref_loss.dynamic_targets(box, gt, C, 0.7)):
        np.testing.assert_array_equal(o.cpu().numpy(), r)


def test_loss_vs_oracle():
    import viddet_b200
    rng = np.random.RandomState(5)
    B, M, C = 3, 20, 20
    gt, box, pre = synth(rng, B, M, C)
    merged = ref_loss.target_merge(box, gt, *pre, C, 0.7)
    N = box.shape[1]
    preds = [rng.normal(0, 2.0, (B, N, w)).astype(np.float32) for w in (1, 2, 2, C)]
    ref = ref_loss.yolo3_loss(*preds, *merged)
    out = viddet_b200.YOLOV3Loss()(*[cuda(t) for t in preds], *[cuda(t) for t in merged])
    for o, r in zip(out, ref):
        np.testing.assert_allclose(o.cpu().numpy(), r, rtol=1e-4)      # fp32 tree reduction vs float64 accumulation
    # deterministic: identical bits on a second run
    out2 = viddet_b200.YOLOV3Loss()(*[cuda(t) for t in preds], *[cuda(t) for t in merged])
    for a, b in zip(out, out2):
        assert torch.equal(a, b)
    # a NaN prediction under a zero weight still poisons the class loss, as in the reference (x*0)
    preds[3][0, 0, 0] = np.nan
    out3 = viddet_b200.YOLOV3Loss()(*[cuda(t) for t in preds], *[cuda(t) for t in merged])
    assert np.isnan(out3[3].cpu().numpy()[0]) and np.isfinite(out3[3].cpu().numpy()[1])


def test_train_errors():
    import viddet_b200
    mg = viddet_b200.YOLOV3TargetMerger(3, 0.7)
    with pytest.raises(viddet_b200.VidDetError):
        mg._run(torch.zeros((1, 4, 4)).cuda(), torch.zeros((1, 2000, 4)).cuda(), None)     # M > 1024
