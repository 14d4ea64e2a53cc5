"""Robustness of the speculative thresholds: ONE session / workspace processes different iid input sets in turn
(the thresholds of a frame slot were left by a different frame each time).  Reports redone frames and step time."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
wl = sys.argv[1] if len(sys.argv) > 1 else "voc416_b64"
C, size, frames = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(99)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
sets = [bench.synth_tips(torch, gen, frames, size, dev) for _ in range(6)]
s = head.session(sets[0])
redone = []
for i in range(24):
    s.rebind(sets[i % 6]); s.run(); redone.append(s.redone_frames())
print("redone frames per call (6 different iid sets in turn, %d frames each):" % frames, redone)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 120
e0.record()
for i in range(n):
    s.rebind(sets[i % 6]); s.run()
e1.record(); torch.cuda.synchronize()
print("%s serial step with fresh inputs every call: %.1f us (direct launches)" % (wl, 1e3 * e0.elapsed_time(e1) / n))
s.rebind(sets[0])
for i in range(4): s.run()
torch.cuda.synchronize()
e0.record()
for i in range(n): s.run()
e1.record(); torch.cuda.synchronize()
print("%s serial step, same inputs every call:        %.1f us (direct launches)" % (wl, 1e3 * e0.elapsed_time(e1) / n))
