"""Fixed cost of the head / nms stages: time vs batch size (graph replay)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200
from viddet_b200 import _lib
import bench
C, size = 20, 416
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
for frames in (1, 4, 16, 32, 64, 128):
    sessions = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(4)]
    for s in sessions: s.run()
    torch.cuda.synchronize()
    out = []
    for stage, name in ((_lib.VD_STAGE_HEAD, "head"), (_lib.VD_STAGE_NMS, "nms"), (_lib.VD_STAGE_ALL, "all")):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(8): sessions[i % 4].run(stage)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        out.append("%s %.1f" % (name, 1e3 * e0.elapsed_time(e1) / 80))
    print("frames %3d: %s us/step" % (frames, "  ".join(out)))
