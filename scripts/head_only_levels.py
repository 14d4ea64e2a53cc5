"""Head kernels only (no NMS stage), back to back in one graph, 4 batches per launch, for the epilogue debug levels
(VD_DEBUG_SKIP_EPILOGUE: 0 full, 1 mainloop only, 7 box part only, 8 no box store, 9 nothing emitted): which part of the epilogue costs what."""
import json
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import viddet_b200
    dev = torch.device("cuda", 0)
    C, size, frames = bench.WORKLOADS["voc416_b64"]
    G = 4
    gen = torch.Generator(device=dev).manual_seed(1234)
    head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
    head.set_nms(0.45, 400, 100)
    pool = [bench.synth_tips(torch, gen, frames * G, size, dev) for _ in range(4)]
    lvl = os.environ.pop("LEVEL")
    sessions = [head.session(pool[j]) for j in range(3)]
    for s in sessions:
        s.run(); s.run()                      # thresholds in place (full calls, level 0)
    torch.cuda.synchronize()
    os.environ["VD_DEBUG_SKIP_EPILOGUE"] = lvl
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16):
            sessions[i % 3].rebind(pool[i % 4]).run(2)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"level": int(lvl), "us_per_step_head_only": e0.elapsed_time(e1) * 1e3 / (160 * G)}))
else:
    for lvl in ("0", "1", "7", "8", "9"):
        env = dict(os.environ, LEVEL=lvl)
        subprocess.run([sys.executable, __file__, "child"], env=env)
