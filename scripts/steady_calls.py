"""Runs N full fused calls (head + NMS) on rotating sessions: the target of `ncu -k regex:... -s <skip> -c 1`."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
wl = sys.argv[1] if len(sys.argv) > 1 else "voc416_b64"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
group = int(sys.argv[3]) if len(sys.argv) > 3 else 1         # batches per launch (bench.py's GROUP)
C, size, frames = bench.WORKLOADS[wl]
frames *= group
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
ss = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(3)]
for i in range(n): ss[i % 3].run()
torch.cuda.synchronize()
print("done")
