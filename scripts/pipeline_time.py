"""Pipelined step time (head of batch j || NMS of batch j-1) vs the serial step, with a result check."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200
from viddet_b200 import _lib
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "voc416_b64"
C, size, frames = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
sessions = [head.session(bench.synth_tips(torch, gen, frames, size, dev), return_keep=True) for _ in range(4)]
ref = []
for s in sessions:
    s.run(); torch.cuda.synchronize()
    ref.append((s.keep.clone(), s.scores.clone()))
    s.keep.fill_(-7); s.scores.fill_(-7.0)
pipe = viddet_b200.HeadPipeline(sessions)
for s in sessions: s.keep.fill_(-7); s.scores.fill_(-7.0)
n = 200
for i in range(2): pipe.cycle()
torch.cuda.synchronize()
for j, s in enumerate(sessions):
    assert torch.equal(s.keep, ref[j][0]) and torch.equal(s.scores, ref[j][1]), "pipelined results differ for session %d" % j
print("pipelined results identical to the serial call")
for s in sessions: s.capture()
for name, fn in (("serial graph", lambda i: sessions[i % 4].replay()), ("pipelined", lambda i: pipe.cycle() if i % pipe.steps_per_cycle == 0 else None)):
    for i in range(8): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / n
    gb = bench.algorithmic_bytes_per_frame(C, size) * frames / (us * 1e-6) / 1e9
    print("%s %s: %.1f us/step  %.0f frames/s (%.0f GB/s algorithmic, %.1f%% of HBM peak)" % (wl, name, us, frames / (us * 1e-6), gb, 100 * gb / 6431.1))
