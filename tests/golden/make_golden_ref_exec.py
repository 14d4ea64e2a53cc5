"""Generates tests/golden/ref_exec_golden.npz by EXECUTING the reference's own classes
  YOLOOutputV3                          models/definitions/yolo/yolo3.py:25-199
  YOLOV3PrefetchTargetGenerator         models/definitions/yolo/yolo_target.py:13-148
  YOLOV3DynamicTargetGeneratorSimple    models/definitions/yolo/yolo_target.py:151-204
  YOLOV3TargetMerger                    models/definitions/yolo/yolo_target.py:207-281
  TimeDistributed, TemporalPooling      models/definitions/layers.py:161-264
  YOLOV3 (+ YOLODetectionBlockV3, _conv2d, _upsample)   yolo3.py:202-534, layers.py:10-20,63-70  -- inference forward after the stages
  YOLOV3Temporal (+ its YOLODetectionBlockV3 / YOLOOutputV3, Conv, _conv3d, _conv21d)   yolo3_temporal.py:25-555, layers.py:73-158
                                                        -- t=5, t_out=True, conv type 21, inference forward after the stages
over tests/golden/mx_shim.py (a numpy stand-in for the MXNet / GluonCV operators they call; MXNet itself cannot be imported here).
The class sources are cut out of /root/reference with `ast` at run time and exec'd -- nothing is copied into the repo.
What this pins: the reference's own logic (slicing, reshape/transposes = row order, the per-GT loop, index math, _slice,
where-merges); what stays restated: the upstream operators inside the shim.  NumPy here is 2.x (NEP 50): np.float32 scalars
mixed with Python numbers stay fp32, whereas the reference's era promoted them to float64 -- the target VALUES in
this file are therefore the 'nep50' variant (<= 1 ulp from the legacy ones); assignments and class rows are identical.
Authoring container only; the .npz travels."""
import ast
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mx_shim  # noqa: E402
from mx_shim import ND, F  # noqa: E402

REF = "/root/reference/models/definitions/yolo"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_exec_golden.npz")
f32 = np.float32


def load_classes(path, names, ns=None):
    """exec the named top-level classes / functions of a reference file into `ns` (a fresh shim namespace by default)."""
    src = open(path).read()
    ns = mx_shim.namespace() if ns is None else ns
    for n in ast.parse(src).body:
        if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names:
            exec(ast.get_source_segment(src, n), ns)
    return [ns[n] for n in names]


def bf16_round(x):
    u = np.ascontiguousarray(x, f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(f32).reshape(np.shape(x))


ANCHORS = [[116, 90, 156, 198, 373, 326], [30, 61, 62, 45, 59, 119], [10, 13, 16, 30, 33, 23]]     # output order s32,s16,s8
STRIDES = [32, 16, 8]


def main():
    (YOLOOutputV3,) = load_classes(os.path.join(REF, "yolo3.py"), ["YOLOOutputV3"])
    Prefetch, Dynamic, Merger = load_classes(os.path.join(REF, "yolo_target.py"),
                                             ["YOLOV3PrefetchTargetGenerator", "YOLOV3DynamicTargetGeneratorSimple", "YOLOV3TargetMerger"])
    # the merger instantiates the dynamic generator by name
    Merger.__init__.__globals__["YOLOV3DynamicTargetGeneratorSimple"] = Dynamic
    rng = np.random.RandomState(20241019)
    out = {}

    # ---------------- YOLOOutputV3: inference, train-mode 7-tuple, agnostic
    dec_cases = [(3, 0, 5, 4, 64, 2), (20, 1, 6, 6, 64, 1), (4, 2, 3, 7, 128, 2)]       # (C, scale index, H, W, Cin, B)
    for ci, (C, si, H, W, Cin, B) in enumerate(dec_cases):
        x = bf16_round(rng.standard_normal((B, Cin, H, W)).astype(f32))
        n = 3 * (5 + C)
        w = bf16_round(rng.uniform(-0.07, 0.07, (n, Cin, 1, 1)).astype(f32))
        b = rng.uniform(-0.3, 0.3, n).astype(f32)
        pre = "dec%d_" % ci
        out.update({pre + "x": x, pre + "w": w, pre + "b": b, pre + "meta": np.array([C, si, H, W, Cin, B])})
        for mode in ("infer", "train", "agnostic"):
            blk = YOLOOutputV3(si, C, ANCHORS[si], STRIDES[si], agnostic=(mode == "agnostic"))
            blk.prediction.weight, blk.prediction.bias = w, b
            mx_shim.autograd.training = (mode == "train")
            res = blk.hybrid_forward(F, ND(x), blk.anchors, blk.offsets)
            mx_shim.autograd.training = False
            if mode == "train":
                for k, r in zip(("bbox", "raw_centers", "raw_scales", "objness", "class_pred", "anchors", "offsets"), res):
                    out[pre + "train_" + k] = r.asnumpy()
            else:
                out[pre + mode] = res.asnumpy()

    # ---------------- YOLOV3PrefetchTargetGenerator (+ merger on top of its outputs)
    tg_cases = [(20, 2, 8, 128, False, False), (20, 3, 12, 416, False, True), (7, 2, 6, 160, True, False)]   # (C, B, M, size, multi-hot, mixup)
    for ci, (C, B, M, size, multi, mix) in enumerate(tg_cases):
        gt = np.full((B, M, 4), -1.0, f32)
        ids = np.zeros((B, M, C), f32) if multi else np.full((B, M, 1), -1.0, f32)
        for b in range(B):
            nb = rng.randint(1, M + 1) if b else M                     # image 0 full, others ragged (-1 padding ends the loop)
            w_ = np.exp(rng.uniform(np.log(4), np.log(size * 0.8), nb)); h_ = np.exp(rng.uniform(np.log(4), np.log(size * 0.8), nb))
            cx = np.floor(rng.uniform(w_ / 2, size - w_ / 2)) + 0.37; cy = np.floor(rng.uniform(h_ / 2, size - h_ / 2)) + 0.61
            gt[b, :nb] = np.stack([np.maximum(cx - w_ / 2, 0), np.maximum(cy - h_ / 2, 0),
                                   np.minimum(cx + w_ / 2, size - 1e-3), np.minimum(cy + h_ / 2, size - 1e-3)], -1)
            if nb >= 2:                                                # two GTs landing on the same cell + anchor: last writer wins
                gt[b, 1] = gt[b, 0] + f32(0.25)
            if multi:
                for m in range(nb):
                    ids[b, m, rng.choice(C, size=rng.randint(1, 4), replace=False)] = 1.0
            else:
                ids[b, :nb, 0] = rng.randint(0, C, nb)
        mixr = rng.uniform(0.3, 1.0, (B, M, 1)).astype(f32) if mix else None
        hw = [size // s for s in STRIDES]
        img = ND(np.zeros((B, 3, size, size), f32))
        xs = [ND(np.zeros((B, 1, h, h), f32)) for h in hw]
        anchors = [ND(np.asarray(a, f32).reshape(1, 1, -1, 2)) for a in ANCHORS]
        offsets = []
        for h in hw:
            gx, gy = np.meshgrid(np.arange(h), np.arange(h))
            offsets.append(ND(np.concatenate((gx[:, :, None], gy[:, :, None]), -1).reshape(1, -1, 1, 2).astype(f32)))
        gen = Prefetch(C)
        res = gen.forward(img, xs, anchors, offsets, ND(gt), ND(ids), ND(mixr) if mix else None)
        pre = "tg%d_" % ci
        out.update({pre + "gt": gt, pre + "ids": ids, pre + "meta": np.array([C, B, M, size, int(multi), int(mix)])})
        if mix:
            out[pre + "mix"] = mixr
        names = ("objectness", "center", "scale", "weights", "class")
        for k, r in zip(names, res):
            out[pre + k] = r.asnumpy()
        # merger: predictions = jittered copies of a few GT boxes scattered over the anchor rows
        N = res[0].shape[1]
        preds = np.zeros((B, N, 4), f32)
        xy = rng.uniform(0, size * 0.8, (B, N, 2)); wh = rng.uniform(8, size * 0.5, (B, N, 2))
        preds[..., :2] = xy; preds[..., 2:] = xy + wh
        for b in range(B):
            rows = rng.choice(N, size=min(N, 40), replace=False)
            src = gt[b, rng.randint(0, max(1, int((gt[b, :, 0] >= 0).sum())), size=len(rows))]
            preds[b, rows] = src + rng.normal(0, 2.0, src.shape).astype(f32)
        merged = Merger(C, 0.7)(ND(preds), ND(gt), *res)
        out[pre + "preds"] = preds
        for k, r in zip(names + ("class_mask",), merged):
            out[pre + "merged_" + k] = r.asnumpy()
    # ---------------- TimeDistributed(YOLOOutputV3) on a (B,T,C,H,W) window, TemporalPooling max / mean
    TemporalPooling, TimeDistributed = load_classes("/root/reference/models/definitions/layers.py", ["TemporalPooling", "TimeDistributed"])
    C, si, H, W, Cin, B, T = 5, 1, 4, 5, 64, 2, 3
    x = bf16_round(rng.standard_normal((B, T, Cin, H, W)).astype(f32))
    n = 3 * (5 + C)
    w = bf16_round(rng.uniform(-0.07, 0.07, (n, Cin, 1, 1)).astype(f32)); b = rng.uniform(-0.3, 0.3, n).astype(f32)
    blk = YOLOOutputV3(si, C, ANCHORS[si], STRIDES[si])
    blk.prediction.weight, blk.prediction.bias = w, b
    blk_call = lambda z: blk.hybrid_forward(F, z, blk.anchors, blk.offsets)
    td = TimeDistributed(blk_call)
    out.update({"td_x": x, "td_w": w, "td_b": b, "td_meta": np.array([C, si, H, W, Cin, B, T]),
                "td_out": td.hybrid_forward(F, ND(x)).asnumpy(),
                "pool_max": TemporalPooling(T, "max").hybrid_forward(F, ND(x)).asnumpy(),
                "pool_mean": TemporalPooling(T, "mean").hybrid_forward(F, ND(x)).asnumpy()})
    # ---------------- YOLOV3.hybrid_forward (inference) after the backbone stages: blocks, transitions, upsample + concat, outputs, NMS
    ns = mx_shim.namespace()
    load_classes("/root/reference/models/definitions/layers.py", ["_upsample", "_conv2d"], ns)
    load_classes(os.path.join(REF, "yolo_target.py"), ["YOLOV3DynamicTargetGeneratorSimple", "YOLOV3TargetMerger"], ns)
    YOLOV3, = load_classes(os.path.join(REF, "yolo3.py"), ["YOLOOutputV3", "YOLODetectionBlockV3", "YOLOV3"], ns)[2:]
    C, B = 4, 2
    hw = [(16, 16), (8, 8), (4, 4)]                                 # s8, s16, s32 maps of a 128 x 128 input
    stage_ch = [64, 128, 192]
    feats = [bf16_round(rng.standard_normal((B, c, h, w)).astype(f32)) for c, (h, w) in zip(stage_ch, hw)]
    feats = [np.where(f > 0, f, 0.1 * f).astype(f32) for f in feats]
    feats = [bf16_round(f) for f in feats]
    stages = [(lambda z, f=f: ND(f)) for f in feats]               # backbone stand-ins: stage i emits the i-th pyramid level
    mx_shim.PARAM_LOG.clear()
    mx_shim.PARAM_RNG.seed(77)
    net = YOLOV3(stages, [128, 128, 128], [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                 [8, 16, 32], classes=["c%d" % i for i in range(C)])
    ids, scores, bboxes = net.hybrid_forward(F, ND(np.zeros((B, 3, 128, 128), f32)))
    net.nms_thresh = 0                                             # yolo3.py:525: NMS skipped -> the plain concat of the decoded rows
    rid, rsc, rbb = net.hybrid_forward(F, ND(np.zeros((B, 3, 128, 128), f32)))
    # the training-mode branch without a recorded loss (yolo3.py:498-509,532-535): raw predictions concatenated over the scales
    net.nms_thresh = 0.45
    mx_shim.autograd.training = True
    raw = net.hybrid_forward(F, ND(np.zeros((B, 3, 128, 128), f32)))
    mx_shim.autograd.training = False
    for k, r in zip(("box_preds", None, None, None, "box_centers", "box_scales", "objness", "class_pred"), raw):
        if k is not None:
            out["neck_train_" + k] = r.asnumpy()
    out["neck_train_fmap_shapes"] = np.array([m.shape for m in raw[3]])
    out["neck_train_offsets_sizes"] = np.array([m.size for m in raw[2]])
    out.update({"neck_ids": ids.asnumpy(), "neck_scores": scores.asnumpy(), "neck_bboxes": bboxes.asnumpy(),
                "neck_det": np.concatenate([rid.asnumpy(), rsc.asnumpy(), rbb.asnumpy()], -1),
                "neck_meta": np.array([C, B, len(mx_shim.PARAM_LOG)])})
    for i, f in enumerate(feats):
        out["neck_feat%d" % i] = f
    # parameters are NOT stored (6 MB): the test replays the shim's draws from RandomState(77) (tests/util.py::replay_shim_params)
    out["neck_param_kinds"] = np.array([0 if e[0] == "conv" else 1 for e in mx_shim.PARAM_LOG])
    out["neck_param_shapes"] = np.array([list(e[1].shape) + [0] * (4 - e[1].ndim) for e in mx_shim.PARAM_LOG])
    out["neck_param_bias"] = np.array([1 if (e[0] == "conv" and e[2] is not None) else 0 for e in mx_shim.PARAM_LOG])
    out["neck_param_check"] = np.array([float(np.asarray(e[1], np.float64).sum()) for e in mx_shim.PARAM_LOG])
    # ---------------- YOLOV3Temporal.hybrid_forward (inference), t=5, t_out=True, conv type 21: TimeDistributed stages / outputs /
    # transitions, (2+1)D detection blocks with swapaxes around the 3-D convs, 5-D upsample + concat, 4-D box_nms and slice
    ns = mx_shim.namespace()
    load_classes("/root/reference/models/definitions/layers.py", ["_upsample", "_conv2d", "_conv3d", "_conv21d", "Conv", "TimeDistributed"], ns)
    load_classes(os.path.join(REF, "yolo_target.py"), ["YOLOV3DynamicTargetGeneratorSimple", "YOLOV3TargetMerger"], ns)
    YOLOV3Temporal, = load_classes(os.path.join(REF, "yolo3_temporal.py"), ["YOLOOutputV3", "YOLODetectionBlockV3", "YOLOV3Temporal"], ns)[2:]
    C, B, T = 3, 1, 5
    hw = [(8, 8), (4, 4), (2, 2)]
    stage_ch = [64, 128, 192]
    tfeats = [np.where((f := rng.standard_normal((B, T, c, h, w)).astype(f32)) > 0, f, 0.1 * f).astype(f32) for c, (h, w) in zip(stage_ch, hw)]
    tfeats = [bf16_round(f) for f in tfeats]
    tstages = [(lambda z, f=f: ND(f.reshape((-1,) + f.shape[2:]))) for f in tfeats]      # called through TimeDistributed: (B*T, C, H, W)
    mx_shim.PARAM_LOG.clear()
    mx_shim.PARAM_RNG.seed(78)
    tnet = YOLOV3Temporal(tstages, [128, 128, 128], [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]],
                          [8, 16, 32], classes=["c%d" % i for i in range(C)], t=5, t_out=True, conv=21)
    tids, tscores, tbboxes = tnet.hybrid_forward(F, ND(np.zeros((B, T, 3, 64, 64), f32)))
    tnet.nms_thresh = 0
    rid, rsc, rbb = tnet.hybrid_forward(F, ND(np.zeros((B, T, 3, 64, 64), f32)))
    out.update({"tneck_ids": tids.asnumpy(), "tneck_scores": tscores.asnumpy(), "tneck_bboxes": tbboxes.asnumpy(),
                "tneck_det": np.concatenate([rid.asnumpy(), rsc.asnumpy(), rbb.asnumpy()], -1),
                "tneck_meta": np.array([C, B, T, len(mx_shim.PARAM_LOG)])})
    for i, f in enumerate(tfeats):
        out["tneck_feat%d" % i] = f
    out["tneck_param_kinds"] = np.array([0 if e[0] == "conv" else 1 for e in mx_shim.PARAM_LOG])
    out["tneck_param_shapes"] = np.array([list(e[1].shape) + [0] * (5 - e[1].ndim) for e in mx_shim.PARAM_LOG])
    out["tneck_param_bias"] = np.array([1 if (e[0] == "conv" and e[2] is not None) else 0 for e in mx_shim.PARAM_LOG])
    out["tneck_param_check"] = np.array([float(np.asarray(e[1], np.float64).sum()) for e in mx_shim.PARAM_LOG])
    out["n_dec"] = np.array(len(dec_cases)); out["n_tg"] = np.array(len(tg_cases))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
