"""oracle/ref_post.py::hierarchical_nms against golden vectors produced by EXECUTING the reference's own
hierarchical_nms / iou / CombinedDetection tree methods (tests/golden/make_golden_hier_nms.py): this row's parity is pinned."""
import os

import numpy as np

from oracle import ref_post

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hier_nms_golden.npz")


def cases():
    g = np.load(GOLD)
    for ci in range(int(g["n_cases"])):
        pre = "c%d_" % ci
        yield ci, {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}


def test_oracle_matches_reference_golden():
    stats = {}
    for ci, c in cases():
        ov, conf, lvl = c["params"]
        out, cnt = ref_post.hierarchical_nms(c["rows"], c["counts"], c["levels"], c["parent"], c["branch"], ov, conf, int(lvl), stats=stats)
        np.testing.assert_array_equal(cnt, c["out_counts"], err_msg="case %d" % ci)
        np.testing.assert_array_equal(out, c["out_rows"], err_msg="case %d" % ci)
    # the goldens exercise every branch of the merge rule
    assert all(stats.get(k, 0) > 0 for k in ("new", "off_branch", "same_cls_max", "ignored")), stats


def test_hand_cases():
    # chain 0 <- 1 <- 2 (levels 1,2,3), class 3 a separate root
    levels = np.array([1, 2, 3, 1]); parent = np.array([-1, 0, 1, -1])
    branch = np.array([[1, 1, 1, 0], [1, 1, 1, 0], [1, 1, 1, 0], [0, 0, 0, 1]], np.uint8)
    a = [10, 10, 50, 50]
    far = [200, 200, 260, 260]
    rows = np.full((1, 4, 6), -1.0, np.float32)
    rows[0, 0] = [0, 0.9] + a            # the parent class, same place as its grandchild: ignored (a child already stands there)
    rows[0, 1] = [2, 0.6] + a            # leaf first (sorted by class desc)
    rows[0, 2] = [3, 0.5] + far          # elsewhere: a new box
    rows[0, 3] = [2, 0.8] + [11, 10, 50, 50]   # same class, IoU > 0.5: confidences max-ed into the first leaf box
    out, cnt = ref_post.hierarchical_nms(rows, np.array([4]), levels, parent, branch)
    assert cnt[0] == 2
    np.testing.assert_allclose(out[0, 0], [3, 0.5] + far)
    np.testing.assert_allclose(out[0, 1], [2, 0.8] + a)
    # level_thresh = 1 lifts every class to its level-1 ancestor before merging
    out, cnt = ref_post.hierarchical_nms(rows, np.array([4]), levels, parent, branch, level_thresh=1)
    assert cnt[0] == 2 and out[0, 1, 0] == 0 and np.isclose(out[0, 1, 1], 0.9)
    # only the best-overlapping kept box is consulted (:771-783): an off-branch box in the same place hides the leaf behind it
    rows[0, 2] = [3, 0.5] + a
    out, cnt = ref_post.hierarchical_nms(rows, np.array([4]), levels, parent, branch)
    assert cnt[0] == 4


def test_class_tree_tables_match_reference_tree_methods():
    """viddet_b200.ClassTree.tables (host logic) reproduces the levels / on_branch tables the reference's CombinedDetection
    methods produced for the golden trees."""
    from viddet_b200.blocks import ClassTree
    for ci, c in cases():
        n = len(c["parent"])
        names = ["n%04d" % i for i in range(n)]
        parents = {names[i]: ("ROOT" if c["parent"][i] < 0 else names[c["parent"][i]]) for i in range(n)}
        levels, parent, branch = ClassTree.tables(names, parents)
        np.testing.assert_array_equal(levels, c["levels"])
        np.testing.assert_array_equal(parent, c["parent"])
        np.testing.assert_array_equal(branch, c["branch"])
