"""First call on a fresh workspace (every frame goes through the exact path) vs the steady state, CUDA events, voc416_b64."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
C, size, frames = bench.WORKLOADS["voc416_b64"]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
warm = head.session(bench.synth_tips(torch, gen, frames, size, dev)); warm.run(); warm.run(); torch.cuda.synchronize()   # module load, attributes
tips = bench.synth_tips(torch, gen, frames, size, dev)
for trial in range(2):
    s = head.session(tips)
    ts = []
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("fresh session, calls 1-4: " + ", ".join("%.0f us" % t for t in ts), " redone frames after call 4:", s.redone_frames())
