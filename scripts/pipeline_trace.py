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
for i in range(12): pipe.step(i)
torch.cuda.synchronize()
hw = [(size // st) ** 2 for st in bench.STRIDES]
anc = 3 * sum(hw); tif = sum((h + 127) // 128 for h in hw); F = frames
def al(x): return (x + 255) // 256 * 256
off = (256 + 256 + al(F * 4096 * 4) + al(F * anc * 16) + al(F * tif * 1024 * 8) + al(F * tif * 4)) // 8
# last step 11: head of session 3, NMS of session 2
hd = ss[3]._ws.view(torch.int64)[off + 4096: off + 4096 + 4 * 148].cpu().view(148, 4).numpy()
nm = ss[2]._ws.view(torch.int64)[off + 8192: off + 8192 + 4 * F].cpu().view(F, 4).numpy()
t0 = min(hd[:, 0].min(), nm[:, 0].min())
print("head CTAs : start %.1f..%.1f  end %.1f..%.1f us" % ((hd[:, 0].min() - t0) / 1e3, (hd[:, 0].max() - t0) / 1e3, (hd[:, 2].min() - t0) / 1e3, (hd[:, 2].max() - t0) / 1e3))
print("NMS  CTAs : start %.1f..%.1f  end %.1f..%.1f us ; duration med %.1f max %.1f" % ((nm[:, 0].min() - t0) / 1e3, (nm[:, 0].max() - t0) / 1e3,
      (nm[:, 1].min() - t0) / 1e3, (nm[:, 1].max() - t0) / 1e3, np.median(nm[:, 1] - nm[:, 0]) / 1e3, (nm[:, 1] - nm[:, 0]).max() / 1e3))
hs = set(hd[:, 3].tolist()); ns = nm[:, 2].tolist()
print("NMS SMs distinct %d, shared with head SMs %d" % (len(set(ns)), len(set(ns) & hs)))
shared = set(ns)
a = hd[[i for i in range(148) if hd[i, 3] in shared]]; b = hd[[i for i in range(148) if hd[i, 3] not in shared]]
print("head CTA duration on SMs with NMS: med %.1f ; without: med %.1f us" % (np.median(a[:, 2] - a[:, 0]) / 1e3 if len(a) else -1, np.median(b[:, 2] - b[:, 0]) / 1e3 if len(b) else -1))
