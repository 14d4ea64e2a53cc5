"""Generates tests/golden/hier_nms_golden.npz by EXECUTING the reference's own `hierarchical_nms` and `iou`
(detect_yolo3.py:712-789) and `CombinedDetection.get_levels / generate_branches / on_branch` (datasets/combined.py:99-150).
Their modules import mxnet at the top and cannot be imported, so the function sources are cut out of the files with `ast`
at run time and exec'd here (nothing is copied into the repo).  Inputs are Python floats (what `load_predictions` yields),
so the arithmetic is plain float64 and independent of the NumPy version.  Authoring container only; the .npz travels."""
import ast
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hier_nms_golden.npz")


def cut(path, names, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    out = {}
    for n in body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            out[n.name] = ast.get_source_segment(src, n)
    assert set(out) == set(names), (names, list(out))
    return out


def load_reference():
    ns = {"tqdm": lambda it, **kw: it}
    for name, code in cut(os.path.join(REF, "detect_yolo3.py"), ["iou", "hierarchical_nms"]).items():
        exec(code, ns)
    meth = cut(os.path.join(REF, "datasets", "combined.py"), ["get_levels", "generate_branches", "on_branch"], cls="CombinedDetection")
    cns = {}
    for name, code in meth.items():
        import textwrap
        exec(textwrap.dedent(code), cns)

    class Tree:
        get_levels = cns["get_levels"]
        generate_branches = cns["generate_branches"]
        on_branch = cns["on_branch"]

        def __init__(self, wn_classes, parents):
            self.wn_classes, self.parents = wn_classes, parents
            self.brances, self.branches_ind = self.generate_branches()

    return ns["hierarchical_nms"], Tree


def make_tree(rng, n):
    """Random forest in the reference's encoding: parents precede their children in wn_classes (as in the combined tree,
    where class ids grow from root to leaf: hierarchical_nms relies on it, :756,:780)."""
    names = ["n%04d" % i for i in range(n)]
    parents = {}
    for i, nm in enumerate(names):
        parents[nm] = "ROOT" if i < 3 or rng.uniform() < 0.08 else names[rng.randint(0, i)]
    return names, parents


def make_boxes(rng, n_img, post, C):
    rows = np.full((n_img, post, 6), -1.0, np.float32)
    counts = np.zeros(n_img, np.int32)
    for f in range(n_img):
        n = int(rng.randint(0, post + 1))
        ncl = max(1, n // 4)
        centers = rng.uniform(0.1, 0.9, size=(ncl, 2))
        which = rng.randint(0, ncl, size=n)
        c = centers[which] + rng.normal(0, 0.01, size=(n, 2))
        wh = np.exp(rng.uniform(np.log(0.05), np.log(0.4), size=(ncl, 2)))[which] * np.exp(rng.normal(0, 0.05, size=(n, 2)))
        b = np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, 1)
        rows[f, :n, 0] = rng.randint(0, C, size=n)
        rows[f, :n, 1] = rng.uniform(0.0, 1.0, size=n)
        rows[f, :n, 2:] = b
        counts[f] = n
    return rows, counts


def main():
    hier_nms, Tree = load_reference()
    rng = np.random.RandomState(20241018)
    out = {}
    cases = [(12, 8, 40, 0.5, 0.0, 10), (40, 16, 100, 0.5, 0.0, 2), (40, 16, 100, 0.3, 0.25, 1), (285, 12, 100, 0.5, 0.0, 3),
             (6, 4, 7, 0.5, 0.0, 10)]
    for ci, (C, n_img, post, ov, conf, lvl) in enumerate(cases):
        names, parents = make_tree(rng, C)
        tree = Tree(names, parents)
        levels = np.array(tree.get_levels(), np.int32)
        parent_idx = np.array([names.index(parents[n]) if parents[n] != "ROOT" else -1 for n in names], np.int32)
        branch = np.array([[1 if tree.on_branch(i, j) else 0 for j in range(C)] for i in range(C)], np.uint8)
        rows, counts = make_boxes(rng, n_img, post, C)
        if ci == 0:                       # pixel-space boxes as well (the `+ 1` of iou() is a pixel convention)
            rows[..., 2:] = np.where(rows[..., 2:] >= 0, np.round(rows[..., 2:] * 416), -1)
        preds = {}
        for f in range(n_img):
            preds["img%03d" % f] = [[int(r[0]), float(r[1])] + [float(v) for v in r[2:]] for r in rows[f, :counts[f]]]
        res = hier_nms(preds, tree, ov_thresh=ov, conf_thresh=conf, level_thresh=lvl)
        orow = np.full((n_img, post, 6), -1.0, np.float64)
        ocnt = np.zeros(n_img, np.int32)
        for f in range(n_img):
            lst = res["img%03d" % f]
            ocnt[f] = len(lst)
            for j, b in enumerate(lst):
                orow[f, j] = b
        assert np.array_equal(orow.astype(np.float32).astype(np.float64), orow)      # outputs are copies of fp32 inputs
        pre = "c%d_" % ci
        out.update({pre + "levels": levels, pre + "parent": parent_idx, pre + "branch": branch, pre + "rows": rows,
                    pre + "counts": counts, pre + "params": np.array([ov, conf, lvl], np.float64),
                    pre + "out_rows": orow.astype(np.float32), pre + "out_counts": ocnt})
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", {k: int(out["c%d_out_counts" % k].sum()) for k in range(len(cases))})


if __name__ == "__main__":
    main()
