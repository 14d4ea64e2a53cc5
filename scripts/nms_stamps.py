"""Reads the per-phase clock64 stamps of block 0 of nms_final_hist_kernel (VD_DEBUG_NMS_STAMPS=1)."""
import os, sys, ctypes, torch
os.environ["VD_DEBUG_NMS_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
from viddet_b200 import _lib
C, size, frames = bench.WORKLOADS["voc416_b64"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
s = head.session(bench.synth_tips(torch, gen, frames, size, dev))
for _ in range(5): s.run()
torch.cuda.synchronize()
# stamps live at the start of the (unused) merge area: find its offset by scanning for plausible clocks
ws = s._ws.view(torch.int64)
names = ["start", "hist scan", "stream lists", "sort", "tail start", "box gather", "diag tiles", "(unused)", "wavefront", "finish"]
# offset of listsA: hints 256 + hist + boxes + lists0 + counts0 (mirrors make_plan)
F = frames; anc = 10647; tif = 30
def al(x): return (x + 255) // 256 * 256
off = 256 + 256 + al(F * 4096 * 4) + al(F * anc * 16) + al(F * tif * 1024 * 8) + al(F * tif * 4) * 2 + al(F * 4) + al(F * 256) + al(F * 2048 * 8) + al(F * 4) + 256 + al(F * 4) + al(F * 4)
st = ws[off // 8: off // 8 + 10].cpu().tolist()
print("clocks:", st)
prev = st[0]
for i in range(1, 10):
    if st[i] < st[0]: continue
    print("%-14s %8d cycles  %6.2f us" % (names[i], st[i] - prev, (st[i] - prev) / 1.9e3)); prev = st[i]
print("total %.2f us" % ((st[9] - st[0]) / 1.9e3))
