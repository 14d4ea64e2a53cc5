"""Per-kernel timing + roofline fractions for the non-headline configs of SURVEY.md 8(d):
cfg 3 (COCO-608 head), cfg 4 (temporal VID head, T=5 windows), cfg 5 (target generation, C=285).
CUDA events on the launching stream, inputs resident in HBM, rotating input sets (> L2).
Prints one JSON line per measurement; `python scripts/bench_configs.py [coco] [vid] [targets] [voc]`.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import viddet_b200
from viddet_b200 import _lib
from tests.util import ANCHORS, make_gt

PEAK_HBM, _ = bench.measured_peaks()
PEAKS = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")) else {}
PEAK_TC = float(PEAKS.get("bf16_tflops_sustained", 1413.6))
dev = torch.device("cuda", 0)


def timeit(fn, n=50, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n          # ms


def event_pairs(sessions, n, stages):
    """Median CUDA-event time of each stage inside real calls (stages of a call run back to back, state stays steady)."""
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(n)]
    for i in range(4):
        sessions[i % len(sessions)].run()
    torch.cuda.synchronize()
    for i in range(n):
        s = sessions[i % len(sessions)]
        evs[i][0].record()
        for k, st in enumerate(stages):
            s.run(st)
            evs[i][k + 1].record()
    torch.cuda.synchronize()
    return [sorted(e[k].elapsed_time(e[k + 1]) for e in evs)[n // 2] for k in range(len(stages))]


def head_cfg(name, obj_bias=None):
    """obj_bias: the 'trained-like' secondary variant of SURVEY 8(d): objectness bias (e.g. -5) => sparse candidates."""
    C, size, frames = bench.WORKLOADS[name]
    gen = torch.Generator(device=dev).manual_seed(1234)
    head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
    if obj_bias is not None:
        for o in head.yolo_outputs:
            bias = o.prediction.bias.clone()
            bias.view(3, 5 + C)[:, 4] = obj_bias
            o.prediction.set_data(o.prediction.weight, bias)
        name = "%s_objbias%g" % (name, obj_bias)
    nrot = 3
    sessions = [head.session(bench.synth_tips(torch, gen, frames, size, dev)) for _ in range(nrot)]
    for s in sessions:
        s.capture()
    pipe = viddet_b200.HeadPipeline(sessions, rotations=4)
    alg = bench.algorithmic_bytes_per_frame(C, size) * frames
    flops = 2.0 * 3 * (5 + C) * sum((size // st) ** 2 * c for st, c in zip(bench.STRIDES, bench.CHANNELS)) * frames
    t_head, t_nms = event_pairs(sessions, 30, [_lib.VD_STAGE_HEAD, _lib.VD_STAGE_NMS])
    t_all = timeit(lambda i: sessions[i % nrot].replay())
    t_pipe = timeit(lambda i: pipe.cycle(), n=10, warm=2) / pipe.steps_per_cycle
    for label, t in (("head_kernel (events, in real calls)", t_head), ("nms_kernel (events, in real calls)", t_nms), ("step (serial graph)", t_all), ("step (pipelined graph)", t_pipe)):
        print(json.dumps({"cfg": name, "what": label, "ms": t, "frames_per_s": frames / (t * 1e-3),
                          "alg_GBps": alg / (t * 1e-3) / 1e9, "hbm_frac": alg / (t * 1e-3) / 1e9 / PEAK_HBM,
                          "tflops": flops / (t * 1e-3) / 1e12, "tensor_frac": flops / (t * 1e-3) / 1e12 / PEAK_TC}))


def vid_temporal(windows=64, T=5, C=30, size=416, nrot=2):
    """cfg 4: per step `windows` windows of T frames: temporal (3,1,1) cell at 3 scales + pred conv + decode + NMS of
    all windows*T frames."""
    gen = torch.Generator(device=dev).manual_seed(1234)
    head = viddet_b200.YOLOV3Head(C, temporal="conv21").initialize(generator=torch.Generator().manual_seed(1234))
    sessions = []
    for _ in range(nrot):
        tips = []
        for c, s in zip(bench.CHANNELS, bench.STRIDES):
            h = size // s
            x = torch.randn((windows * T, c, h, h), generator=gen, device=dev)
            x = torch.where(x > 0, x, 0.1 * x).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            tips.append(x.reshape(windows, T, c, h, h))
        sessions.append(head.session(tips))
    hw_c2 = sum((size // s) ** 2 * c * c for s, c in zip(bench.STRIDES, bench.CHANNELS))
    hw_c = sum((size // s) ** 2 * c for s, c in zip(bench.STRIDES, bench.CHANNELS))
    f_tconv = 2.0 * (3 * T - 2) * hw_c2 * windows
    f_pred = 2.0 * 3 * (5 + C) * hw_c * T * windows
    t_tc, t_head, t_nms = event_pairs(sessions, 10, [_lib.VD_STAGE_TCONV, _lib.VD_STAGE_HEAD, _lib.VD_STAGE_NMS])
    t_all = timeit(lambda i: sessions[i % nrot].run(), n=10, warm=2)
    print(json.dumps({"cfg": "vid416_T5_w%d" % windows, "what": "temporal_conv x3", "ms": t_tc,
                      "tflops": f_tconv / (t_tc * 1e-3) / 1e12, "tensor_frac": f_tconv / (t_tc * 1e-3) / 1e12 / PEAK_TC}))
    print(json.dumps({"cfg": "vid416_T5_w%d" % windows, "what": "head_kernel", "ms": t_head,
                      "alg_GBps": hw_c * 2 * T * windows / (t_head * 1e-3) / 1e9}))
    print(json.dumps({"cfg": "vid416_T5_w%d" % windows, "what": "nms_kernel", "ms": t_nms}))
    for s_ in sessions:
        s_.capture()
    pipe = viddet_b200.HeadPipeline(sessions, rotations=4)
    t_pipe = timeit(lambda i: pipe.cycle(), n=6, warm=2) / pipe.steps_per_cycle
    for label, t in (("step", t_all), ("step (pipelined graph: NMS of batch j under the tip cells of batch j+1)", t_pipe)):
        print(json.dumps({"cfg": "vid416_T5_w%d" % windows, "what": label, "ms": t, "windows_per_s": windows / (t * 1e-3),
                          "frames_per_s": windows * T / (t * 1e-3),
                          "tflops": (f_tconv + f_pred) / (t * 1e-3) / 1e12,
                          "tensor_frac": (f_tconv + f_pred) / (t * 1e-3) / 1e12 / PEAK_TC}))


def targets(B=128, M=100, C=285, size=416, multi_hot=True):
    rng = np.random.RandomState(1234)
    gt, ids = make_gt(rng, B, M, size=size, num_class=C, multi_hot=multi_hot)
    gb, gi = torch.from_numpy(gt).to(dev), torch.from_numpy(ids).to(dev)
    gen = viddet_b200.YOLOV3PrefetchTargetGenerator(C)
    hs = [size // s for s in bench.STRIDES]
    xs = [(B, 1, h, h) for h in hs]
    anchors = [np.asarray(a, np.float32).reshape(1, 1, 3, 2) for a in ANCHORS]
    offsets = [np.zeros((1, h * h, 1, 2), np.float32) for h in hs]
    img = (B, 3, size, size)
    n_anch = 3 * sum(h * h for h in hs)
    alg = n_anch * (7 + C) * 4 * B
    t = timeit(lambda i: gen(img, xs, anchors, offsets, gb, gi), n=20, warm=3)
    print(json.dumps({"cfg": "targets_B%d_M%d_C%d_%s" % (B, M, C, "multihot" if multi_hot else "ids"), "what": "fill+scatter (incl. torch.empty + ctypes call)",
                      "ms": t, "images_per_s": B / (t * 1e-3), "alg_GBps": alg / (t * 1e-3) / 1e9,
                      "hbm_frac": alg / (t * 1e-3) / 1e9 / PEAK_HBM}))


def train_side(B=128, M=100, C=285, size=416):
    """8f row 1: target merger + loss forward on the cfg-5 shapes; algorithmic bytes = tensors read + written once."""
    n_anch = 3 * sum((size // s) ** 2 for s in bench.STRIDES)
    g = torch.Generator(device=dev).manual_seed(7)
    box = torch.rand((B, n_anch, 4), generator=g, device=dev) * 300
    box[..., 2:] += box[..., :2] + 8
    gt = torch.rand((B, M, 4), generator=g, device=dev) * 300
    gt[..., 2:] += gt[..., :2] + 8
    obj_t = (torch.rand((B, n_anch, 1), generator=g, device=dev) > 0.995).float()
    pre = [obj_t, torch.rand((B, n_anch, 2), generator=g, device=dev), torch.rand((B, n_anch, 2), generator=g, device=dev),
           torch.rand((B, n_anch, 2), generator=g, device=dev), (torch.rand((B, n_anch, C), generator=g, device=dev) > 0.99).float()]
    mg = viddet_b200.YOLOV3TargetMerger(C, 0.7)
    t = timeit(lambda i: mg(box, gt, *pre), n=10, warm=2)
    alg = B * n_anch * 4 * (4 + 7 + (7 + 2 * C))                  # box_preds + narrow prefetched targets read (class rows only where positive), six outputs written
    print(json.dumps({"cfg": "merge_B%d_M%d_C%d" % (B, M, C), "what": "target_merge (incl. torch.empty + ctypes)", "ms": t,
                      "images_per_s": B / (t * 1e-3), "alg_GBps": alg / (t * 1e-3) / 1e9, "hbm_frac": alg / (t * 1e-3) / 1e9 / PEAK_HBM}))
    merged = mg(box, gt, *pre)
    preds = [torch.randn((B, n_anch, w), generator=g, device=dev) for w in (1, 2, 2, C)]
    loss = viddet_b200.YOLOV3Loss()
    t = timeit(lambda i: loss(*preds, *merged), n=10, warm=2)
    alg = B * n_anch * 4 * ((5 + C) + (7 + 2 * C))
    print(json.dumps({"cfg": "loss_B%d_C%d" % (B, C), "what": "yolo3_loss forward", "ms": t, "images_per_s": B / (t * 1e-3),
                      "alg_GBps": alg / (t * 1e-3) / 1e9, "hbm_frac": alg / (t * 1e-3) / 1e9 / PEAK_HBM}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["coco", "vid", "targets"]      # also: "voc", "vidt" (temporal only)
    print(json.dumps({"peaks": {"hbm_GBps": PEAK_HBM, "bf16_tflops_sustained": PEAK_TC}}))
    if "voc" in which:
        head_cfg("voc416_b64")
    if "voc_trained" in which:
        head_cfg("voc416_b64", obj_bias=-5.0)
    if "coco" in which:
        head_cfg("coco608_b64")
    if "vid" in which:
        head_cfg("vid416_b64")
        vid_temporal()
    if "vidt" in which:
        vid_temporal()
    if "targets" in which:
        targets()
        targets(C=20, multi_hot=False)
    if "train" in which:
        train_side()
