"""box_nms on the GPU vs the oracle: outputs and keep-indices (MXNet's `record`) BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import ref_nms
from tests.util import random_dets

pytestmark = pytest.mark.gpu


def run(d, **kw):
    import viddet_b200
    out, rec = viddet_b200.box_nms(torch.from_numpy(d).cuda(), return_record=True, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy(), rec.cpu().numpy()


def check(d, **kw):
    o_ref, r_ref = ref_nms.box_nms(d, return_record=True, **kw)
    o, r = run(d, **kw)
    np.testing.assert_array_equal(r, r_ref)
    np.testing.assert_array_equal(o.view(np.uint32), o_ref.view(np.uint32))     # bit-exact incl. -1 fill


def test_doc_examples():
    x = np.array([[0, 0.5, 0.1, 0.1, 0.2, 0.2], [1, 0.4, 0.1, 0.1, 0.2, 0.2],
                  [0, 0.3, 0.1, 0.1, 0.14, 0.14], [2, 0.6, 0.5, 0.5, 0.7, 0.8]], np.float32)
    for force in (True, False):
        check(x, overlap_thresh=0.1, coord_start=2, score_index=1, id_index=0, force_suppress=force)


def test_upstream_assumption_kats():
    """oracle/ASSUMPTIONS.md A1-A10 on the CUDA operator: hand-derived keep records, and bit-equality with the oracle."""
    from tests.kat_nms import case_arrays
    for name, d, kw, rec in case_arrays():
        o, r = run(d, **kw)
        np.testing.assert_array_equal(r, rec, err_msg=name)
        check(d, **kw)


@pytest.mark.parametrize("N,topk", [(37, -1), (300, 100), (1000, 400), (5000, 400), (40000, 400), (20000, 1024), (600, 512)])
@pytest.mark.parametrize("force", [False, True])
def test_random_parity(N, topk, force):
    rng = np.random.RandomState(N + topk + force)
    d = random_dets(rng, 3, N, num_class=7, tie_frac=0.2)
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0, score_index=1, coord_start=2, force_suppress=force)


def test_massive_ties_and_all_equal_scores():
    rng = np.random.RandomState(7)
    d = random_dets(rng, 2, 30000, num_class=4, tie_frac=1.0)      # scores on a 1/8 grid
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    d[..., 1] = 0.25                                              # every score identical: order = row index
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)


def test_edge_cases():
    rng = np.random.RandomState(11)
    d = random_dets(rng, 2, 500, num_class=3)
    check(d, overlap_thresh=0.45, valid_thresh=2.0, topk=400, id_index=0)            # all filtered -> all -1
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=1, id_index=0)             # k = 1
    check(d, overlap_thresh=0.0, valid_thresh=0.01, topk=400, id_index=0)            # everything overlapping dies
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=-1)          # no ids -> class-agnostic
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, background_id=1)
    d2 = d.copy(); d2[0, :, 1] = np.nan                                               # a batch with only NaN scores
    check(d2, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    d3 = d.copy(); d3[..., 1] *= -1                                                   # negative scores, valid_thresh < 0
    check(d3, overlap_thresh=0.45, valid_thresh=-10.0, topk=50, id_index=0)
    d4 = d.copy(); d4[..., 4:6] = d4[..., 2:4]                                        # zero-area boxes (0/0 IoU)
    check(d4, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)


def test_4d_batch_and_wide_rows_and_formats():
    rng = np.random.RandomState(13)
    d = random_dets(rng, 6, 800, num_class=5).reshape(2, 3, 800, 6)                  # (B,T,rows,6) like yolo3_temporal.py:545
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0)
    wide = np.concatenate([rng.standard_normal((2, 400, 2)).astype(np.float32), random_dets(rng, 2, 400, 4)], -1)
    check(wide, overlap_thresh=0.5, valid_thresh=0.05, topk=100, coord_start=4, score_index=3, id_index=2)
    c = random_dets(rng, 2, 300, 3)
    ctr = c.copy()
    ctr[..., 2] = (c[..., 2] + c[..., 4]) / 2; ctr[..., 3] = (c[..., 3] + c[..., 5]) / 2
    ctr[..., 4] = c[..., 4] - c[..., 2]; ctr[..., 5] = c[..., 5] - c[..., 3]
    check(ctr, overlap_thresh=0.45, valid_thresh=0.01, topk=200, id_index=0, in_format="center", out_format="center")
    check(ctr, overlap_thresh=0.45, valid_thresh=0.01, topk=200, id_index=0, in_format="center", out_format="corner")
    check(c, overlap_thresh=0.45, valid_thresh=0.01, topk=200, id_index=0, in_format="corner", out_format="center")


def test_full_size_voc_rows_properties():
    """BASELINE size (212 940 rows/image): oracle parity on one image + idempotence on the kept set."""
    import viddet_b200
    rng = np.random.RandomState(17)
    d = random_dets(rng, 2, 212940, num_class=20, invalid_frac=0.0)
    d[..., 1] = (0.2 + 0.2 * rng.uniform(size=d.shape[:2])).astype(np.float32)        # everything valid, like random init
    kw = dict(overlap_thresh=0.45, valid_thresh=0.01, topk=400, id_index=0, score_index=1, coord_start=2)
    check(d, **kw)
    o, r = run(d, **kw)
    kept = o[:, :400]
    o2, r2 = run(np.ascontiguousarray(kept), **kw)                                      # NMS of survivors keeps all of them
    n = (r[:, :400] >= 0).sum(1)
    for b in range(2):
        np.testing.assert_array_equal(o2[b, :n[b]], kept[b, :n[b]])
        assert (np.diff(kept[b, :n[b], 1]) <= 0).all()                                  # sorted by score


@pytest.mark.parametrize("N,topk,B", [(1500, -1, 2), (3000, -1, 3), (5000, 2000, 2), (2049, 0, 1), (9000, 4096, 1)])
@pytest.mark.parametrize("force", [False, True])
def test_more_than_max_topk_candidates(N, topk, B, force):
    """MXNet's default topk = -1 (or a large explicit topk) on inputs longer than VD_MAX_TOPK rows: the general path (global
    bitonic sort + workspace-resident wavefront NMS) must give the oracle's rows and records bit for bit, ties included."""
    rng = np.random.RandomState(N + (topk if topk > 0 else 7) + force)
    d = random_dets(rng, B, N, num_class=5, tie_frac=0.2)
    check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0, score_index=1, coord_start=2, force_suppress=force)
    if not force:
        check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0, background_id=2)
        check(d, overlap_thresh=0.6, valid_thresh=-1.0, topk=topk, id_index=-1)                    # no id column, everything valid
        d[..., 1] = 0.25                                                                            # all scores tie: order = row index
        check(d, overlap_thresh=0.45, valid_thresh=0.01, topk=topk, id_index=0)


def test_errors():
    import viddet_b200
    with pytest.raises(viddet_b200.VidDetError):
        viddet_b200.box_nms(torch.zeros((1, 10, 6)))    # CPU tensor: no fallback
