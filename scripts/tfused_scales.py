"""Per-scale timing of the fused temporal head kernel (cfg 4 shapes: 64 windows of T=5 over a resident clip), CUDA events around the
HEAD stage of real calls.  VD_TFUSED_SCALES selects the scale, VD_TFUSED_DBG=1 skips the decode / filter epilogue."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, viddet_b200
from viddet_b200 import _lib
dev = torch.device("cuda", 0)
W, T, C, size = 64, 5, 30, 416
gen = torch.Generator(device=dev).manual_seed(3)
head = viddet_b200.YOLOV3Head(C, temporal="conv21").initialize(generator=torch.Generator().manual_seed(1234))
head.set_nms(0.45, 400, 100)
clips = []
for i in range(2):
    big = [torch.empty((W + T - 1, c, size // s, size // s), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last) for c, s in zip(bench.CHANNELS, bench.STRIDES)]
    bench.synth_tips(torch, gen, W + T - 1, size, dev, out=big)
    clips.append(big)
sess = [head.session([viddet_b200.ClipWindows(t, 0, W, T) for t in clips[i]]) for i in range(2)]
assert sess[0].fused_tip
for s_ in sess:
    s_.run(); s_.run()
torch.cuda.synchronize()
def time_stage(n=10):
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(n)]
    for i in range(n):
        s_ = sess[i % 2]
        evs[i][0].record(); s_.run(_lib.VD_STAGE_HEAD); evs[i][1].record()
        s_.run(_lib.VD_STAGE_NMS)
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[n // 2]
hw_c2 = [(size // s) ** 2 * c * c for s, c in zip(bench.STRIDES, bench.CHANNELS)]
hw_c = [(size // s) ** 2 * c for s, c in zip(bench.STRIDES, bench.CHANNELS)]
for dbg in ("0", "1"):
    os.environ["VD_TFUSED_DBG"] = dbg
    tot = 0.0
    for s in range(3):
        os.environ["VD_TFUSED_SCALES"] = str(1 << s)
        ms = time_stage()
        fl = (2.0 * 13 * hw_c2[s] + 2.0 * 3 * (5 + C) * hw_c[s] * T) * W
        tot += ms
        print(json.dumps({"scale": s, "dbg": dbg, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}))
    os.environ["VD_TFUSED_SCALES"] = "7"
    print(json.dumps({"dbg": dbg, "sum_ms": round(tot, 4), "all_scales_ms": round(time_stage(), 4)}))
