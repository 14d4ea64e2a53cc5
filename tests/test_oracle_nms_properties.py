"""Size-independent properties of the box_nms oracle (hypothesis-driven): idempotence, sortedness, class-aware separation of the
survivors, compaction with -1 fill, agreement of the C and the numpy restatements."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ref_nms
from tests.util import random_dets

KW = dict(valid_thresh=0.01, id_index=0, score_index=1, coord_start=2)


def iou(a, b):
    w = max(0.0, min(a[2], b[2]) - max(a[0], b[0])); h = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    i = w * h
    u = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - i
    return 0.0 if u <= 0 else i / u


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 160), ncls=st.integers(1, 4), topk=st.sampled_from([-1, 1, 7, 50, 400]),
       thr=st.sampled_from([0.3, 0.45, 0.7]), ties=st.sampled_from([0.0, 0.5]))
def test_box_nms_properties(seed, n, ncls, topk, thr, ties):
    rng = np.random.RandomState(seed)
    d = random_dets(rng, 2, n, num_class=ncls, tie_frac=ties)
    out, rec = ref_nms.box_nms(d, overlap_thresh=thr, topk=topk, return_record=True, **KW)
    assert out.shape == d.shape
    for b in range(2):
        kept = out[b][out[b, :, 0] >= 0]
        m = len(kept)
        assert (out[b, m:] == -1).all()                                        # compaction, -1 fill behind the survivors
        assert (np.diff(kept[:, 1]) <= 0).all()                                # score-descending
        assert (kept[:, 1] > 0.01).all()                                       # strict valid threshold
        if topk > 0:
            assert m <= topk
        for i in range(m):                                                     # survivors of one class do not overlap above the threshold
            for j in range(i):
                if kept[i, 0] == kept[j, 0]:
                    assert iou(kept[i, 2:], kept[j, 2:]) <= thr + 1e-6
        rows = rec[b][rec[b] >= 0]
        np.testing.assert_array_equal(d[b][rows], kept)                        # the record output indexes the input rows
    again = ref_nms.box_nms(out, overlap_thresh=thr, topk=topk, **KW)
    np.testing.assert_array_equal(again, out)                                  # idempotent


def test_c_and_numpy_restatements_agree():
    rng = np.random.RandomState(3)
    for n, topk in [(50, -1), (300, 100), (1000, 400)]:
        d = random_dets(rng, 3, n, num_class=5, tie_frac=0.3)
        a = ref_nms.box_nms(d, overlap_thresh=0.45, topk=topk, **KW)
        b, _ = ref_nms.box_nms_py(d, overlap_thresh=0.45, topk=topk, **KW)
        np.testing.assert_array_equal(a, b)
