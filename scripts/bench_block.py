"""Timing of the conv-BN-LReLU cells of YOLODetectionBlockV3 at the 416^2 scales, B frames.
CUDA events on the launching stream, rotating input sets (> L2 when B is large).  One JSON line per cell:
useful FLOPs = 2*taps*pixels*Cin*Cout (padding taps included, as cuDNN would count them) / time vs the measured
sustained bf16 tensor peak; `mma_row_util` = share of the 128 MMA rows that are real pixels.
`python scripts/bench_block.py [B=64] [T=1]`"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200
from viddet_b200 import _lib

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
PEAK_TC = float(PEAKS.get("bf16_tflops_sustained", 1413.6))
NECK = "neck" in sys.argv
if NECK:
    sys.argv.remove("neck")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1


def timeit(fn, n=20, warm=3, reps=5):
    """n calls captured in one CUDA graph (no host launch overhead between the kernels), replayed `reps` times."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps)


def timeit_eager(fn, n=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def box(H, W, F):
    bw, bh, bf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.load().vd_conv_tile_box(H, W, F, ctypes.byref(bw), ctypes.byref(bh), ctypes.byref(bf)))
    return bw.value, bh.value, bf.value


def main():
    g = torch.Generator().manual_seed(0)
    total_ms, total_fl = 0.0, 0.0
    for ch, cin, hw in ((512, 1024, 13), (256, 768, 26), (128, 384, 52)):
        cells = [("reduce1x1", cin, ch, (1, 1, 1)), ("expand3x3", ch, 2 * ch, (1, 3, 3)), ("reduce1x1b", 2 * ch, ch, (1, 1, 1))]
        if T > 1:
            cells.append(("temporal3x1x1", 2 * ch, 2 * ch, (3, 1, 1)))
        for name, ci, co, k in cells:
            cell = viddet_b200.ConvBNLReLU(ci, co, k).initialize(generator=g)
            xs = [torch.randn(B * T, ci, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
                  .reshape(B, T, ci, hw, hw) for _ in range(3)]
            ms = timeit(lambda i: cell(xs[i % 3]))
            taps = k[0] * k[1] * k[2]
            fl = 2.0 * taps * B * T * hw * hw * ci * co
            F = B * T if k[0] == 1 else T
            bw, bh, bf = (128, 1, 1) if taps == 1 else box(hw, hw, F)
            util = 1.0 if taps == 1 else hw * hw * F / (-(-hw // bw) * -(-hw // bh) * -(-F // bf) * 128.0)
            print(json.dumps({"cell": name, "map": hw, "Cin": ci, "Cout": co, "kernel": k, "frames": B * T, "ms": round(ms, 4),
                              "tflops": round(fl / ms / 1e9, 1), "frac_tensor_peak": round(fl / ms / 1e9 / PEAK_TC, 3),
                              "box": [bw, bh, bf], "mma_row_util": round(util, 3)}))
            # the block runs reduce + expand twice, reduce once more, then the tip's expand
            reps = {"reduce1x1": 1, "expand3x3": 3, "reduce1x1b": 2, "temporal3x1x1": 3}[name]
            total_ms += reps * ms; total_fl += reps * fl
    print(json.dumps({"block_total_ms_est": round(total_ms, 3), "tflops": round(total_fl / total_ms / 1e9, 1),
                      "frac_tensor_peak": round(total_fl / total_ms / 1e9 / PEAK_TC, 3), "frames": B * T}))
    # whole blocks, real chaining
    for conv_type in (["2"] if T == 1 else ["21"]):
        for ch, cin, hw in ((512, 1024, 13), (256, 768, 26), (128, 384, 52)):
            blk = viddet_b200.YOLODetectionBlockV3(ch, conv_type, in_channels=cin).initialize(generator=g)
            shape = (B, cin, hw, hw) if conv_type == "2" else (B, T, cin, hw, hw)
            n = B * (T if conv_type != "2" else 1)
            x = torch.randn(n, cin, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).reshape(shape)
            ms = timeit(lambda i: blk(x))
            fl = sum(2.0 * c.kernel[0] * c.kernel[1] * c.kernel[2] * n * hw * hw * c.in_channels * c.channels for c in blk.cells())
            print(json.dumps({"block": conv_type, "map": hw, "channel": ch, "frames": n, "ms": round(ms, 4),
                              "tflops": round(fl / ms / 1e9, 1), "frac_tensor_peak": round(fl / ms / 1e9 / PEAK_TC, 3)}))


def neck():
    """Backbone routes -> detections (YOLOV3.hybrid_forward after the stages) at VOC-416: 19 conv cells + 2 glue kernels + fused head."""
    g = torch.Generator().manual_seed(1)
    nk = viddet_b200.YOLOV3Neck(20).initialize(generator=g)
    routes = [torch.randn(B, c, 416 // s, 416 // s, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
              for c, s in zip((256, 512, 1024), (8, 16, 32))]
    for _ in range(3):
        nk(routes)                                   # steady state of the head's speculative thresholds
    ms = timeit(lambda i: nk(routes), n=5, warm=2)
    cells = [c for b in nk.yolo_blocks for c in b.cells()] + list(nk.transitions)
    hw = {1024: 13, 768: 26, 384: 52}
    fl = 0.0
    for blk, h in zip(nk.yolo_blocks, (13, 26, 52)):
        fl += sum(2.0 * c.kernel[0] * c.kernel[1] * c.kernel[2] * B * h * h * c.in_channels * c.channels for c in blk.cells())
    fl += 2.0 * B * (13 * 13 * 512 * 256 + 26 * 26 * 256 * 128)                       # transitions
    fl += 2.0 * B * 75 * (13 * 13 * 1024 + 26 * 26 * 512 + 52 * 52 * 256)             # prediction convs
    print(json.dumps({"neck": "routes -> detections (20 calls captured into one graph by the benchmark)", "frames": B, "ms": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1),
                      "tflops": round(fl / ms / 1e9, 1), "frac_tensor_peak": round(fl / ms / 1e9 / PEAK_TC, 3),
                      "gflop_per_frame": round(fl / B / 1e9, 2)}))
    sess = nk.session(routes)                        # the same forward as ONE CUDA graph
    ms = timeit_eager(lambda i: sess.replay(), n=20, warm=3)
    ms_eager = timeit_eager(lambda i: nk(routes), n=20, warm=3)
    print(json.dumps({"neck": "routes -> detections (product API called eagerly: 22 launches + per-call torch.empty, host-bound)", "frames": B, "ms": round(ms_eager, 4),
                      "frames_per_s": round(B / ms_eager * 1e3, 1)}))
    print(json.dumps({"neck": "routes -> detections (NeckSession: one CUDA graph)", "frames": B, "ms": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1),
                      "tflops": round(fl / ms / 1e9, 1), "frac_tensor_peak": round(fl / ms / 1e9 / PEAK_TC, 3)}))


if __name__ == "__main__":
    if NECK:
        neck()
    else:
        main()
