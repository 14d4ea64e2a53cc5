"""Per-tile clock64 stamps of CTA 0 of the fused head kernel (VD_DEBUG_HEAD_STAMPS=1)."""
import os, sys, torch
os.environ["VD_DEBUG_HEAD_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
from viddet_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "voc416_b64"
C, size, frames = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
ss = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(3)]
for i in range(9): ss[i % 3].run()          # full calls: the workspace state (bound, hints) is the steady-state one
torch.cuda.synchronize()
s = ss[2]
hw = [(size // st) ** 2 for st in bench.STRIDES]
anc = 3 * sum(hw); tif = sum((h + 127) // 128 for h in hw); F = frames
def al(x): return (x + 255) // 256 * 256
off = 256 + 256 + al(F * 4096 * 4) + al(F * anc * 16) + al(F * tif * 1024 * 8) + al(F * tif * 4) * 2 + al(F * 4) + al(F * 256) + al(F * 2048 * 8) + al(F * 4) + 256 + al(F * 4) + al(F * 4)
ntile = (F * tif + 147) // 148
st = s._ws.view(torch.int64)[off // 8: off // 8 + 16 * ntile].cpu().view(ntile, 16)
t0 = int(st[0, 0])
order = [0, 1, 2, 3, 4, 8, 9, 10, 5, 13, 6, 14, 7]
names = {0: "mma_wait", 1: "mma_go", 2: "mma_commit", 3: "epi_wait", 4: "epi_ready", 8: "box+bound", 9: "setfast", 10: "sweep", 5: "bar", 
         13: "scored", 6: "sum2+rel", 14: "flushed", 7: "end"}
print("tile " + " ".join("%8s" % names[o] for o in order) + "   (us since first MMA wait; epilogue columns after epi_ready are deltas in ns)")
for i in range(ntile):
    v = [(int(x) - t0) / 1.965e3 for x in st[i]]
    out = []
    prev = None
    for o in order:
        if o in (0, 1, 2, 3, 4): out.append("%8.2f" % v[o]); prev = v[o] if o == 4 else prev
        else: out.append("%8.0f" % ((v[o] - prev) * 1e3)); prev = v[o]
    print("%4d " % i + " ".join(out))

ct = s._ws.view(torch.int64)[off // 8 + 4096: off // 8 + 4096 + 4 * 148].cpu().view(148, 4)
g0 = int(ct[:, 0].min())
import numpy as np
a = (ct[:, :3].numpy() - g0) / 1e3
print("CTA start  us: min %.2f med %.2f max %.2f" % (a[:, 0].min(), np.median(a[:, 0]), a[:, 0].max()))
print("CTA setup  us: min %.2f med %.2f max %.2f" % ((a[:, 1] - a[:, 0]).min(), np.median(a[:, 1] - a[:, 0]), (a[:, 1] - a[:, 0]).max()))
print("CTA end    us: min %.2f med %.2f max %.2f" % (a[:, 2].min(), np.median(a[:, 2]), a[:, 2].max()))
print("ends sorted:", " ".join("%.1f" % x for x in sorted(a[:, 2])[::8]))
print("cta0: start %.2f setup-done %.2f end %.2f" % tuple(a[0]))
