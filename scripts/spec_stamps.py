"""Per-tile clock64 stamps of CTA 0 of the speculative head kernel (VD_DEBUG_HEAD_STAMPS=1), one launch over `group` batches:
where does a tile's time go -- waiting for the accumulator (mainloop-bound) or inside the epilogue (box part / class loop / tail)?"""
import ctypes
import os
import sys

import numpy as np
import torch

os.environ["VD_DEBUG_HEAD_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import viddet_b200
from viddet_b200 import _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "voc416_b64"
group = int(sys.argv[2]) if len(sys.argv) > 2 else 4
C, size, frames = bench.WORKLOADS[wl]
frames *= group
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
ss = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(2)]
for i in range(6):
    ss[i % 2].run()
torch.cuda.synchronize()
s = ss[1]
off = _lib.load().vd_head_debug_offset(ctypes.byref(s.params))
st = s._ws.view(torch.int64)[off // 8: off // 8 + 16 * 250].cpu().numpy().reshape(250, 16).astype(np.float64)
G = 3
used = [i for i in range(250) if st[i, 4] > 0 and st[i, 7] > 0]
t0 = st[used[0], 0]
us = lambda c: c / 1965.0
rows = []
print("tile grp | mma_wait mma_go mma_commit | epi_tileknown acc_ready box_done class_done end  (us since the first MMA wait) | acc_wait box class tail total (us)")
for i in used:
    v = st[i]
    rows.append((us(v[4] - v[3]), us(v[8] - v[4]), us(v[10] - v[8]), us(v[7] - v[10]), us(v[7] - v[3]), us(v[2] - v[1])))
    if i < 40 or i % 10 == 0:
        print("%4d %3d | %8.2f %8.2f %8.2f | %8.2f %8.2f %8.2f %8.2f %8.2f | %6.2f %6.2f %6.2f %6.2f %6.2f" % (
            i, i % G, us(v[0] - t0), us(v[1] - t0), us(v[2] - t0), us(v[3] - t0), us(v[4] - t0), us(v[8] - t0), us(v[10] - t0), us(v[7] - t0),
            rows[-1][0], rows[-1][1], rows[-1][2], rows[-1][3], rows[-1][4]))
r = np.array(rows)
print("tiles stamped: %d of CTA 0" % len(rows))
for k, name in enumerate(["wait for the accumulator", "box part", "class loop", "tail (release + reservation)", "whole epilogue turn", "MMA issue -> commit"]):
    print("%-30s mean %.2f us  median %.2f  p90 %.2f" % (name, r[:, k].mean(), np.median(r[:, k]), np.quantile(r[:, k], 0.9)))
# gaps in the MMA stream: time the MMA warp waited for a free accumulator
w = [us(st[i, 1] - st[i, 0]) for i in used]
print("MMA warp waiting for a free accumulator: mean %.2f us per tile, total %.1f us of %.1f us" % (np.mean(w), np.sum(w), us(st[used[-1], 2] - t0)))
