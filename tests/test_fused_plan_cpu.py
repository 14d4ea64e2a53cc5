"""Host-side work plan of the fused temporal head kernel (csrc/tfused.cuh::tfused_schedule, through vd_head_fused_tip_plan): for any
shape and pair count the CTA pairs' strided rounds + contiguous ranges cover every item of every scale exactly once, and the modelled
load is balanced.  No device work (runs on the CPU-only box)."""
import ctypes

import numpy as np
import pytest

import viddet_b200
from viddet_b200 import _lib

CHANNELS = [1024, 512, 256]


def plan(windows, T, size, pairs, channels=CHANNELS):
    p = _lib.VdHeadParams()
    p.num_scales, p.num_class, p.frames, p.T, p.K_frames, p.join = 3, 30, windows * T, T, 1, 0
    p.nms_thresh, p.valid_thresh, p.nms_topk, p.post_nms = 0.45, 0.01, 400, 100
    for i, (c, st) in enumerate(zip(channels, (32, 16, 8))):
        s = p.scale[i]
        s.H = s.W = size // st
        s.Cin = c
        s.stride = float(st)
        s.tip_nhwc_bf16 = 16            # plan only: non-null, aligned dummies
        s.weight_bf16 = 16
    items = (ctypes.c_int * 3)()
    strided = (ctypes.c_int * 3)()
    beg = (ctypes.c_int * (3 * 81))()
    rc = viddet_b200.load().vd_head_fused_tip_plan(ctypes.byref(p), pairs, items, strided, beg)
    assert rc == 0, viddet_b200.load().vd_last_error()
    return list(items), list(strided), np.array(list(beg)).reshape(3, 81), p


@pytest.mark.parametrize("windows,T,size,pairs", [(64, 5, 416, 74), (64, 5, 416, 72), (3, 5, 160, 74), (8, 5, 256, 74), (1, 5, 416, 74),
                                                   (7, 3, 320, 66), (200, 5, 416, 74), (2, 1, 96, 5), (33, 5, 608, 80)])
def test_every_item_exactly_once_and_balanced(windows, T, size, pairs):
    items, strided, beg, p = plan(windows, T, size, pairs)
    loads = np.zeros(pairs)
    for s in range(3):
        hw = (size // (32 >> s)) ** 2
        m_tiles = -(-T * hw // 128)
        assert items[s] == (windows * m_tiles + 1) // 2
        seen = np.zeros(items[s], dtype=np.int64)
        cost = (CHANNELS[s] // 256) * (0.7 * 3 * CHANNELS[s] / 64 + 1.5)
        assert strided[s] * pairs <= items[s]
        assert beg[s, 0] == strided[s] * pairs and beg[s, pairs] == items[s]
        for c in range(pairs):
            mine = [r * pairs + c for r in range(strided[s])] + list(range(beg[s, c], beg[s, c + 1]))
            assert beg[s, c] <= beg[s, c + 1]
            for it in mine:
                seen[it] += 1
            loads[c] += cost * len(mine)
        assert (seen == 1).all(), (s, np.flatnonzero(seen != 1)[:5])
        if hw < 8 * 128:
            assert strided[s] == 0                   # strided rounds only where the +-HW neighbour rows lie many tiles apart
    # longest-processing-time greedy: no pair carries more than the lightest pair plus one item of the costliest scale that has items
    biggest = max((CHANNELS[s] // 256) * (0.7 * 3 * CHANNELS[s] / 64 + 1.5) for s in range(3) if items[s])
    assert loads.max() - loads.min() <= biggest + 1e-6, (loads.max(), loads.min(), biggest)


def test_fused_tip_flag_and_applicability():
    _, _, _, p = plan(64, 5, 416, 74)
    lib = viddet_b200.load()
    for i in range(3):                                  # the kernel needs the tip cell's parameters and scratch
        s = p.scale[i]
        s.tconv_weight_bf16 = s.tconv_scale = s.tconv_shift = s.tconv_out_nhwc_bf16 = 16
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 1
    p.flags = _lib.VD_HEAD_NO_FUSED_TIP
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 0
    p.flags = 0
    p.num_class = 80                                    # no fused shape for COCO: separate kernels
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 0
    p.num_class = 20
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 1
    p.num_class = 23                                    # padded onto the 30-class shape
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 1
    p.num_class = 31                                    # two class windows of the 80-class shape: separate kernels
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 0
    p.num_class = 20
    p.scale[2].Cin = 128                                # channel counts must be multiples of 256
    assert lib.vd_head_fused_tip(ctypes.byref(p)) == 0
