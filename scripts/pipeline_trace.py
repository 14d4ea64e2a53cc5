"""Where and when did the head CTAs of batch j and the NMS CTAs of batch j-1 run? (VD_DEBUG_HEAD_STAMPS)"""
import os, sys, torch, numpy as np
os.environ["VD_DEBUG_HEAD_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
from viddet_b200 import _lib
C, size, frames = bench.WORKLOADS["voc416_b64"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
ss = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(4)]
pipe = viddet_b200.HeadPipeline(ss)
for i in range(3): pipe.cycle()
torch.cuda.synchronize()
hw = [(size // st) ** 2 for st in bench.STRIDES]
anc = 3 * sum(hw); tif = sum((h + 127) // 128 for h in hw); F = frames
def al(x): return (x + 255) // 256 * 256
off = (256 + 256 + al(F * 4096 * 4) + al(F * anc * 16) + al(F * tif * 1024 * 8) + al(F * tif * 4) * 2 + al(F * 4) + al(F * 256) + al(F * 2048 * 8) + al(F * 4) + 256 + al(F * 4) + al(F * 4)) // 8
# consecutive pipelined steps: absolute windows
for i in range(2): pipe.cycle()
torch.cuda.synchronize()
w = []
for j in (0, 1, 2, 3):
    hd = ss[j]._ws.view(torch.int64)[off + 4096: off + 4096 + 4 * 148].cpu().view(148, 4).numpy()
    nm = ss[j]._ws.view(torch.int64)[off + 8192: off + 8192 + 4 * F].cpu().view(F, 4).numpy()
    w.append((hd[:, 0].min(), hd[:, 2].max(), nm[:, 0].min(), nm[:, 1].max()))
t0 = w[0][0]
for j in range(4):
    print("step %d: head(%d) %.1f..%.1f (end stamp is a lower bound: special-register reads are not ordered by barriers) | nms(%d) %.1f..%.1f" % (j, j, (w[j][0] - t0) / 1e3, (w[j][1] - t0) / 1e3, j, (w[j][2] - t0) / 1e3, (w[j][3] - t0) / 1e3))

