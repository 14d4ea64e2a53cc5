"""Where does a pipelined step's time go?  (a) a graph of n steps launched from an idle GPU (what `bench.py --steps n` times),
(b) the same graph replayed back to back (launch latency hidden), for several n; fresh inputs vs each session replaying its own
batch.  `VD_LIB=<variant.so> python scripts/graph_latency.py` compares build variants (e.g. -DVD_FALLBACK_CTAS=8)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import viddet_b200

dev = torch.device("cuda", 0)
C, size, frames = bench.WORKLOADS["voc416_b64"]
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
head.set_nms(0.45, 400, 100)
pool = [bench.synth_tips(torch, gen, frames, size, dev) for _ in range(16)]
sessions = [head.session(pool[j]) for j in range(bench.NRING)]


def med(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def measure(pipe, n, reps=10):
    for _ in range(3):
        pipe.cycle()
    torch.cuda.synchronize()
    idle = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); pipe.cycle(); e1.record()
        torch.cuda.synchronize()
        idle.append(e0.elapsed_time(e1) * 1e3 / n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        pipe.cycle()
    e1.record()
    torch.cuda.synchronize()
    return med(idle), min(idle), e0.elapsed_time(e1) * 1e3 / (n * reps)


for mode in ("fresh", "same"):
    for n in (4, 8, 20, 64):
        pipe = viddet_b200.HeadPipeline(sessions, steps=n, inputs=pool if mode == "fresh" else None)
        a, amin, b = measure(pipe, n)
        print(json.dumps({"lib": os.path.basename(viddet_b200.SO_PATH), "inputs": mode, "steps_per_graph": n, "us_per_step_from_idle_median": a,
                          "us_per_step_from_idle_min": amin, "us_per_step_back_to_back": b}))
        del pipe
# head kernels alone (no NMS stage): the main stream's lower bound
for s in sessions:
    s.run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(64):
        sessions[i % 4].rebind(pool[i % 16]).run(2)
for _ in range(2):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"what": "head kernels only, back to back (64 per graph)", "us_per_step": e0.elapsed_time(e1) * 1e3 / 640}))
