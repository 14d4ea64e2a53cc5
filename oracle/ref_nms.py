"""Oracle: MXNet `contrib.box_nms` / `contrib.box_iou` (ctypes wrapper over ref_nms.c, plus an
independent pure-numpy restatement used to cross-check the C on small cases).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned by the reference
(the operators live in un-vendored MXNet); algorithm = SURVEY.md Appendix A.3.
Reference call sites: yolo3.py:526-528, yolo3_temporal.py:545-547, yolo_target.py:92.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libref_nms.so")
_lib = None
_FMT = {"corner": 0, "center": 1}


def build(force: bool = False) -> str:
    """Compile ref_nms.c (gcc) -> oracle/_build/libref_nms.so."""
    src = os.path.join(_HERE, "ref_nms.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lib.ref_box_nms.restype = ctypes.c_int
        lib.ref_box_nms.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, fp, ip]
        lib.ref_box_iou.restype = ctypes.c_int
        lib.ref_box_iou.argtypes = [fp, ctypes.c_int64, fp, ctypes.c_int64, fp]
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def box_nms(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1,
            id_index=-1, background_id=-1, force_suppress=False, in_format="corner",
            out_format="corner", return_record=False, threads=1):
    """mx.nd.contrib.box_nms.  data (..., num_elem, width>=6) fp32 -> same shape (+ record)."""
    lib = _load()
    data = np.ascontiguousarray(data, dtype=np.float32)
    shape = data.shape
    assert data.ndim >= 2
    n_elem, width = shape[-2], shape[-1]
    nb = int(np.prod(shape[:-2])) if data.ndim > 2 else 1
    flat = data.reshape(nb, n_elem, width)
    out = np.empty_like(flat)
    rec = np.empty((nb, n_elem), dtype=np.int32)

    def run(lo, hi):
        if hi <= lo:
            return
        lib.ref_box_nms(_fp(flat[lo:hi]), hi - lo, n_elem, width, overlap_thresh, valid_thresh,
                        int(topk), coord_start, score_index, id_index, background_id,
                        int(bool(force_suppress)), _FMT[in_format], _FMT[out_format],
                        _fp(out[lo:hi]), rec[lo:hi].ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))

    threads = max(1, min(int(threads), nb))
    if threads == 1:
        run(0, nb)
    else:   # images are independent (MXNet parallelises its CPU kernels the same way)
        cuts = np.linspace(0, nb, threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda se: run(*se), zip(cuts[:-1], cuts[1:])))
    out = out.reshape(shape)
    if return_record:
        return out, rec.reshape(shape[:-1])
    return out


def box_iou(lhs, rhs):
    """mx.nd.contrib.box_iou(lhs (...,4), rhs (...,4), format='corner') -> lhs.shape[:-1]+rhs.shape[:-1]."""
    lib = _load()
    lhs = np.ascontiguousarray(lhs, dtype=np.float32)
    rhs = np.ascontiguousarray(rhs, dtype=np.float32)
    n = lhs.size // 4
    m = rhs.size // 4
    out = np.empty((n, m), dtype=np.float32)
    lib.ref_box_iou(_fp(lhs), n, _fp(rhs), m, _fp(out))
    return out.reshape(lhs.shape[:-1] + rhs.shape[:-1])


# ---------------------------------------------------------------------------------------------
# independent restatement (pure numpy / python loops) -- small cases only
# ---------------------------------------------------------------------------------------------
def box_nms_py(data, overlap_thresh=0.5, valid_thresh=0.0, topk=-1, coord_start=2, score_index=1,
               id_index=-1, background_id=-1, force_suppress=False):
    """Literal Appendix-A.3 steps 1-6 with numpy fp32 scalars (corner format only)."""
    f32 = np.float32
    data = np.asarray(data, dtype=f32)
    shape = data.shape
    n_elem, width = shape[-2], shape[-1]
    flat = data.reshape(-1, n_elem, width)
    out = np.full_like(flat, -1.0)
    rec = np.full(flat.shape[:2], -1, dtype=np.int32)
    k_eff = min(topk, n_elem) if topk > 0 else n_elem
    for b in range(flat.shape[0]):
        d = flat[b]
        valid = [i for i in range(n_elem) if d[i, score_index] > f32(valid_thresh)
                 and not (id_index >= 0 and background_id >= 0 and int(d[i, id_index]) == background_id)]
        order = sorted(valid, key=lambda i: -float(d[i, score_index]))   # python sort is stable
        order = order[:k_eff]
        box = d[:, coord_start:coord_start + 4]
        def _area(i):          # BoxArea(): 0 for a negative extent (oracle/ASSUMPTIONS.md A3)
            w, h = f32(box[i, 2] - box[i, 0]), f32(box[i, 3] - box[i, 1])
            return f32(0) if (w < 0 or h < 0) else f32(w * h)
        area = {i: _area(i) for i in order}
        dead = set()
        for ri, r in enumerate(order):
            if r in dead:
                continue
            for p in order[ri + 1:]:
                if p in dead:
                    continue
                if (not force_suppress) and id_index >= 0 and int(d[r, id_index]) != int(d[p, id_index]):
                    continue
                iw = f32(min(box[r, 2], box[p, 2]) - max(box[r, 0], box[p, 0]))
                ih = f32(min(box[r, 3], box[p, 3]) - max(box[r, 1], box[p, 1]))
                iw = iw if iw > 0 else f32(0)
                ih = ih if ih > 0 else f32(0)
                inter = f32(iw * ih)
                with np.errstate(divide="ignore", invalid="ignore"):
                    iou = f32(inter / f32(f32(area[r] + area[p]) - inter))
                if iou > f32(overlap_thresh):
                    dead.add(p)
        row = 0
        for r in order:
            if r in dead:
                continue
            out[b, row] = d[r]
            rec[b, row] = r
            row += 1
    return out.reshape(shape), rec.reshape(shape[:-1])
