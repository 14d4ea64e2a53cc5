"""Times n-step CUDA graphs (sessions chained in one graph) to separate launch overhead from kernel time."""
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
sessions = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(4)]
for s in sessions: s.run()
torch.cuda.synchronize()
for stage, name in ((_lib.VD_STAGE_HEAD, "head"), (_lib.VD_STAGE_NMS, "nms"), (_lib.VD_STAGE_ALL, "all")):
    for nsteps in (1, 4, 16):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(nsteps): sessions[i % 4].run(stage)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(4, 128 // nsteps)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / (reps * nsteps)
        print("%s %s: %d-step graph: %.1f us/step" % (wl, name, nsteps, us))
