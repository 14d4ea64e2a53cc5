"""Times only the fused head kernel (VD_STAGE_HEAD) for a workload; prints us per 64-frame batch."""
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
stages = ((_lib.VD_STAGE_HEAD, "head"),) if os.environ.get("VD_DEBUG_SKIP_EPILOGUE") else ((_lib.VD_STAGE_HEAD, "head"), (_lib.VD_STAGE_NMS, "nms"), (_lib.VD_STAGE_ALL, "all"))
for s_ in sessions: s_.run()
for stage, name in stages:
    for s_ in sessions: s_.capture(stage)
    for i in range(8): sessions[i % 4].replay(stage)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 100
    e0.record()
    for i in range(n): sessions[i % 4].replay(stage)
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / n
    gb = bench.algorithmic_bytes_per_frame(C, size) * frames / (us * 1e-6) / 1e9
    print("%s %s: %.1f us/step  (%.0f GB/s algorithmic)" % (wl, name, us, gb))
