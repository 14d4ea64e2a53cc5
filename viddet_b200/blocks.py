"""Host-side mirror of the reference's block / operator surface for the detection-head hot path.

Same names, argument meaning and return layouts as HaydenFaulkner/VidDet (MXNet/Gluon), on torch
CUDA tensors used purely as carriers: every function below is a thin ctypes call into one C-ABI
symbol of libviddet_b200.so (include/viddet_b200.h).  No CPU path, no torch math on the hot path.

reference surface                                             here
---------------------------------------------------------------------------------------------
mx.nd.contrib.box_nms        (yolo3.py:526-528)               box_nms()
YOLOOutputV3                 (yolo3.py:25-199)                YOLOOutputV3
TimeDistributed              (layers.py:208-264)              TimeDistributed
TemporalPooling              (layers.py:161-205)              TemporalPooling
Conv('21') temporal cell     (layers.py:82-89)                TemporalTipConv
_conv2d / _conv3d cells      (layers.py:63-79)                ConvBNLReLU
YOLODetectionBlockV3         (yolo3_temporal.py:184-239)      YOLODetectionBlockV3
YOLOV3.hybrid_forward after the stages (yolo3.py:496-534)     YOLOV3Neck, upsample_concat
YOLOV3 / YOLOV3Temporal tail (yolo3.py:496,522-556;           YOLOV3Head
                              yolo3_temporal.py:468,542-555)
YOLOV3PrefetchTargetGenerator(yolo_target.py:13-148)          YOLOV3PrefetchTargetGenerator
"""
from __future__ import annotations

import ctypes
import warnings

import numpy as np
import torch

from . import _lib
from ._lib import VdHeadParams, check, load, ptr, stream_ptr

# wrappers.py:80-84 lists (s8,s16,s32); yolo3.py:416-417 reverses -> output order s32,s16,s8
DEFAULT_ANCHORS = [[116, 90, 156, 198, 373, 326], [30, 61, 62, 45, 59, 119], [10, 13, 16, 30, 33, 23]]
DEFAULT_STRIDES = [32, 16, 8]
DEFAULT_CHANNELS = [1024, 512, 256]
_FMT = {"corner": 0, "center": 1}


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.VidDetError(-1, "%s must be a CUDA tensor (viddet_b200 has no CPU path)" % name)


# ------------------------------------------------------------------------------------------------
# layout carriers
# ------------------------------------------------------------------------------------------------
def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) tensor in any layout/dtype -> bf16 tensor of the same logical shape whose memory is
    channels-last (B,H,W,C).  fp32 NCHW (the reference's layout) goes through the repack kernel."""
    _require_cuda(x, "x")
    assert x.dim() == 4, "expected (B,C,H,W)"
    if x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last):
        return x
    if x.dtype == torch.float32 and x.is_contiguous():
        B, C, H, W = x.shape
        out = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        check(load().vd_repack_nchw_f32_to_nhwc_bf16(ptr(x), ptr(out), B, C, H, W, stream_ptr()))
        return out
    return x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


class SplitF32:
    """fp32 feature map carried for the tensor cores as P bf16 planes (fp32-parity modes, include/viddet_b200.h):
    `data` (P, B, H, W, C) bf16, plane-major, p0 = bf16(v), p1 = bf16(v - p0), p2 = bf16(v - p0 - p1); `shape` is the
    logical (B, C, H, W)."""

    def __init__(self, data, shape):
        self.data, self.shape = data, tuple(shape)
        self.planes = int(data.shape[0])

    @property
    def device(self):
        return self.data.device


def to_nhwc_split(x, planes=3) -> SplitF32:
    """(B,C,H,W) fp32 NCHW (the reference's layout and dtype) -> SplitF32 through the repack kernel."""
    if isinstance(x, SplitF32):
        assert x.planes == planes, "carrier has %d planes, the call needs %d" % (x.planes, planes)
        return x
    _require_cuda(x, "x")
    assert x.dim() == 4, "expected (B,C,H,W)"
    x = x.to(torch.float32).contiguous()
    B, C, H, W = x.shape
    out = torch.empty((planes, B, H, W, C), dtype=torch.bfloat16, device=x.device)
    check(load().vd_repack_nchw_f32_to_nhwc_split(ptr(x), ptr(out), B, C, H, W, planes, stream_ptr()))
    return SplitF32(out, (B, C, H, W))


def window_frame_indices(num_frames, centre, window_size, window_step=1):
    """Frame indices of the temporal window around frame `centre` of a clip of `num_frames` frames, exactly as the reference's
    dataset builds it (datasets/imgnetvid.py:480-506): floor(window_size / 2) frames back and forward at `window_step`, indices
    clamped to the clip (the first / last frame is repeated at the clip's ends), an even window drops its last frame."""
    half = int(window_size / 2.0)
    win = [max(0, centre - back) for back in range(half * window_step, window_step - 1, -window_step)]
    win.append(centre)
    for fwd in range(window_step, half * window_step + 1, window_step):
        if len(win) == window_size:
            break
        win.append(min(num_frames - 1, centre + fwd))
    return win


class ClipWindows:
    """`count` windows of `T` consecutive frames sliding frame by frame over a RESIDENT clip (one scale's features of a whole
    clip, (L, C, H, W) channels-last bf16): window b = clip frames [start + b, start + b + T).  That is what
    window_frame_indices gives for consecutive centres start + T//2 + b away from the clip's ends (window_step 1); the temporal
    head reads the windows straight out of the clip through one overlapping TMA map (vd_temporal_conv_ex, window stride 1) --
    no frame is copied, every frame is fetched from HBM once instead of T times.  The clamped windows at a clip's two ends repeat
    frames: build those with `materialise_windows`."""

    def __init__(self, clip, start, count, T):
        clip = to_nhwc_bf16(clip)
        L, C, H, W = clip.shape
        assert 0 <= start and count >= 0 and start + count + T - 1 <= L, "windows leave the clip"
        self.clip, self.start, self.count, self.T = clip, int(start), int(count), int(T)
        self.shape = (self.count * self.T, C, H, W)          # frames the head sees
        self.device = clip.device

    def data_ptr(self):
        return self.clip[self.start:].data_ptr()

    def materialise(self):
        """The same windows as an ordinary (count, T, C, H, W) tensor (T-fold copy; for tests and the parity modes)."""
        idx = torch.arange(self.count, device=self.device)[:, None] + torch.arange(self.T, device=self.device)[None, :] + self.start
        return self.clip[idx.reshape(-1)].reshape((self.count, self.T) + tuple(self.clip.shape[1:]))


def materialise_windows(clip, centres, window_size, window_step=1):
    """(len(centres), T, C, H, W) windows gathered out of a resident clip with the reference's clamping (window_frame_indices):
    the general path -- any centres, any step, the clip's ends."""
    clip = to_nhwc_bf16(clip)
    idx = [window_frame_indices(clip.shape[0], int(c), window_size, window_step) for c in centres]
    T = len(idx[0])
    flat = torch.as_tensor(idx, device=clip.device).reshape(-1)
    return clip[flat].reshape((len(idx), T) + tuple(clip.shape[1:]))


def _precision_code(precision):
    if precision is None or precision == "bf16" or (precision == _lib.VD_PREC_BF16 and precision is not False):
        return _lib.VD_PREC_BF16
    if precision in ("fp32", "f32", "float32") or (precision == _lib.VD_PREC_FP32_SPLIT and precision is not True):
        return _lib.VD_PREC_FP32_SPLIT
    if precision == "bf16x2" or precision == _lib.VD_PREC_BF16X2:
        return _lib.VD_PREC_BF16X2
    raise ValueError("precision must be 'bf16', 'bf16x2' or 'fp32', got %r" % (precision,))


# ------------------------------------------------------------------------------------------------
# box_nms
# ------------------------------------------------------------------------------------------------
_WS_CACHE = {}


def _workspace(nbytes, device, owner="nms"):
    """Cached scratch buffer per (device, owner, current stream).  The fused head keeps state in its workspace between
    calls (tile-scheduler counters, histograms left zeroed, thresholds), so it never shares a buffer with box_nms, with
    another head instance (`owner` carries the instance id) or with a call enqueued on another stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), owner,
           torch.cuda.current_stream(device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


def box_nms(data, overlap_thresh=0.5, valid_thresh=0, topk=-1, coord_start=2, score_index=1, id_index=-1,
            background_id=-1, force_suppress=False, in_format="corner", out_format="corner",
            return_record=False):
    """mx.nd.contrib.box_nms (signature and semantics of the MXNet operator).

    data (..., num_elem, width) fp32 -> tensor of the same shape; with return_record also the
    operator's hidden second output: original row index of every kept element, -1 elsewhere."""
    _require_cuda(data, "data")
    if data.dtype != torch.float32:
        raise _lib.VidDetError(-1, "box_nms: data must be float32")
    if in_format not in _FMT or out_format not in _FMT:
        raise _lib.VidDetError(-1, "box_nms: format must be 'corner' or 'center'")
    assert data.dim() >= 2
    d = data.contiguous()
    n_elem, width = d.shape[-2], d.shape[-1]
    nb = 1
    for s in d.shape[:-2]:
        nb *= s
    out = torch.empty_like(d)
    rec = torch.empty(d.shape[:-1], dtype=torch.int32, device=d.device) if return_record else None
    lib = load()
    ws = _workspace(lib.vd_box_nms_workspace_bytes(nb, n_elem, width, int(topk)), d.device)
    check(lib.vd_box_nms(ptr(d), nb, n_elem, width, float(overlap_thresh), float(valid_thresh), int(topk),
                         int(coord_start), int(score_index), int(id_index), int(background_id),
                         int(bool(force_suppress)), _FMT[in_format], _FMT[out_format], ptr(out), ptr(rec),
                         ptr(ws), ws.numel(), stream_ptr()))
    return (out, rec) if return_record else out


# ------------------------------------------------------------------------------------------------
# YOLOOutputV3
# ------------------------------------------------------------------------------------------------
class _Conv1x1:
    """Parameter holder standing in for `nn.Conv2D(all_pred, kernel_size=1)` (yolo3.py:62)."""

    def __init__(self):
        self.weight = None      # (N, Cin, 1, 1) fp32 master copy, Gluon layout
        self.bias = None        # (N,) fp32
        self._w_bf16 = None

    def set_data(self, weight, bias=None):
        weight = torch.as_tensor(weight)
        assert weight.dim() == 4 and weight.shape[2] == 1 and weight.shape[3] == 1
        self.weight = weight.detach().to(torch.float32).cuda().contiguous()
        n = self.weight.shape[0]
        self.bias = (torch.zeros(n, device="cuda") if bias is None
                     else torch.as_tensor(bias).detach().to(torch.float32).cuda().contiguous())
        self._w_bf16 = self.weight.reshape(n, -1).to(torch.bfloat16).contiguous()
        self._w_split = {}

    @property
    def weight_bf16(self):
        return self._w_bf16

    def weight_split(self, planes):
        """(N, planes, Cin) bf16 planes of the fp32 master weight (carrier of the fp32-parity modes), built on first use."""
        if planes not in self._w_split:
            n = self.weight.shape[0]
            w2 = self.weight.reshape(n, -1).contiguous()
            out = torch.empty((n, planes, w2.shape[1]), dtype=torch.bfloat16, device=w2.device)
            check(load().vd_split_f32_rows(ptr(w2), ptr(out), n, w2.shape[1], planes, stream_ptr()))
            self._w_split[planes] = out
        return self._w_split[planes]


class YOLOOutputV3:
    """YOLO output layer V3 (yolo3.py:25-199): 1x1 prediction conv + decode.

    Parameters as in the reference: index, num_class, anchors, stride, alloc_size, agnostic.
    (`k`/`rnn_shape`, the ConvRNN option at yolo3.py:58-60, is out of scope.)
    `__call__(x)` takes (B,Cin,H,W) [fp32 NCHW like the reference, or bf16 channels-last] and returns
      inference: (B, C*H*W*A, 6) rows [id, score, x1, y1, x2, y2]     (yolo3.py:191-197)
      agnostic:  (B, H*W*A, 6)                                           (yolo3.py:184-188)
      training=True: (bbox (B,HW*A,4), raw_centers (B,HW,A,2), raw_scales (B,HW,A,2),
                      objness (B,HW,A,1), class_pred (B,HW,A,C), anchors, offsets)   (yolo3.py:179-182)
    """

    def __init__(self, index, num_class, anchors, stride, alloc_size=(128, 128), k=None, rnn_shape=None,
                 k_join_type="max", agnostic=False, in_channels=None, precision="bf16"):
        self._precision = _precision_code(precision)     # 'fp32': hi/lo split operands, 1e-5 vs the fp32 reference (fp32 NCHW inputs)
        if k is not None and rnn_shape is not None:
            raise NotImplementedError("ConvRNN prediction (yolo3.py:58-60) is outside the hot path")
        anchors = np.array(anchors).astype("float32")
        self._index = index
        self._classes = num_class
        self._num_pred = 1 + 4 + num_class
        self._num_anchors = anchors.size // 2
        self._stride = stride
        self._agnostic = agnostic
        self._alloc_size = tuple(alloc_size)
        self.prediction = _Conv1x1()
        self._anchors_np = anchors.reshape(-1)
        self.anchors = torch.from_numpy(anchors.reshape(1, 1, -1, 2)).cuda()
        gx, gy = np.meshgrid(np.arange(alloc_size[1]), np.arange(alloc_size[0]))
        off = np.concatenate((gx[:, :, None], gy[:, :, None]), axis=-1)[None, None].astype("float32")
        self.offsets = torch.from_numpy(off).cuda()
        if in_channels is not None:
            self.initialize(in_channels)

    # -- parameters ------------------------------------------------------------------------------
    def initialize(self, in_channels, scale=0.07, generator=None):
        """net.initialize() default: weight ~ U(-0.07, 0.07), bias 0 (detect_yolo3.py:885)."""
        n = self._num_pred * self._num_anchors
        w = (torch.rand((n, in_channels, 1, 1), generator=generator) * 2 - 1) * scale
        self.prediction.set_data(w, torch.zeros(n))
        return self

    def reset_class(self, classes, reuse_weights=None):
        """yolo3.py:76-129: new predictor for len(classes) classes, optionally re-using rows."""
        old_classes, old_num_pred = self._classes, self._num_pred
        old_w, old_b = self.prediction.weight, self.prediction.bias
        self._classes = len(classes)
        self._num_pred = 1 + 4 + len(classes)
        in_channels = old_w.shape[1]
        self.initialize(in_channels)
        if reuse_weights:
            assert isinstance(reuse_weights, dict)
            new_w, new_b = self.prediction.weight.clone(), self.prediction.bias.clone()
            for k, v in reuse_weights.items():
                if k >= self._classes or v >= old_classes:
                    warnings.warn("reuse mapping {}/{} -> {}/{} out of range".format(k, self._classes, v, old_classes))
                    continue
                for i in range(self._num_anchors):
                    off_new, off_old = i * self._num_pred, i * old_num_pred
                    new_w[1 + 4 + k + off_new] = old_w[1 + 4 + v + off_old]
                    new_b[1 + 4 + k + off_new] = old_b[1 + 4 + v + off_old]
                    new_w[off_new:1 + 4 + off_new] = old_w[off_old:1 + 4 + off_old]
                    new_b[off_new:1 + 4 + off_new] = old_b[off_old:1 + 4 + off_old]
            self.prediction.set_data(new_w, new_b)

    # -- forward ---------------------------------------------------------------------------------
    def predict(self, x):
        """The prediction conv alone: (B,Cin,H,W) -> pred (B, A*(5+C), H, W) fp32 (yolo3.py:157)."""
        planes = _lib.PLANES[self._precision]
        split = planes > 1
        xb = to_nhwc_split(x, planes) if split else to_nhwc_bf16(x)
        B, Cin, H, W = xb.shape
        n = self._num_pred * self._num_anchors
        if self.prediction.weight is None or self.prediction.weight.shape[1] != Cin:
            raise _lib.VidDetError(-1, "YOLOOutputV3: prediction weights not set for Cin=%d" % Cin)
        pred = torch.empty((B, n, H, W), dtype=torch.float32, device=xb.device)
        if split:
            check(load().vd_pred_conv_ex(ptr(xb.data), B, H, W, Cin, 1, _lib.VD_JOIN_NONE, self._precision,
                                         ptr(self.prediction.weight_split(planes)), ptr(self.prediction.bias), n, ptr(pred), stream_ptr()))
        else:
            check(load().vd_pred_conv(ptr(xb), B, H, W, Cin, 1, _lib.VD_JOIN_NONE, ptr(self.prediction.weight_bf16),
                                      ptr(self.prediction.bias), n, ptr(pred), stream_ptr()))
        return pred

    def decode(self, pred, training=False, out=None, rows_total=None, row_offset=0):
        """yolo3.py:158-199 on a conv output `pred` (B, A*(5+C), H, W) fp32."""
        _require_cuda(pred, "pred")
        pred = pred.contiguous()
        B, N, H, W = pred.shape
        A, C = self._num_anchors, self._classes
        assert N == A * (5 + C)
        if H > self._alloc_size[0] or W > self._alloc_size[1]:
            raise _lib.VidDetError(-1, "feature map larger than alloc_size")
        anc = (ctypes.c_float * (2 * A))(*self._anchors_np.tolist())
        lib, dev = load(), pred.device
        if training:
            bbox = torch.empty((B, H * W * A, 4), device=dev)
            rc = torch.empty((B, H * W, A, 2), device=dev)
            rs = torch.empty((B, H * W, A, 2), device=dev)
            ob = torch.empty((B, H * W, A, 1), device=dev)
            cp = torch.empty((B, H * W, A, C), device=dev)
            check(lib.vd_yolo_decode(ptr(pred), B, H, W, C, A, anc, float(self._stride), _lib.VD_MODE_TRAIN,
                                     ptr(bbox), H * W * A, 0, ptr(rc), ptr(rs), ptr(ob), ptr(cp), stream_ptr()))
            offsets = self.offsets[:, :, :H, :W, :].reshape(1, -1, 1, 2)
            return bbox, rc, rs, ob, cp, self.anchors, offsets
        mode = _lib.VD_MODE_AGNOSTIC if self._agnostic else _lib.VD_MODE_INFER
        rows = H * W * A * (1 if self._agnostic else C)
        if out is None:
            out = torch.empty((B, rows, 6), device=dev)
            rows_total, row_offset = rows, 0
        check(lib.vd_yolo_decode(ptr(pred), B, H, W, C, A, anc, float(self._stride), mode, ptr(out),
                                 rows_total, row_offset, None, None, None, None, stream_ptr()))
        return out

    def __call__(self, x, training=False):
        return self.decode(self.predict(x), training=training)

    hybrid_forward = __call__


# ------------------------------------------------------------------------------------------------
# temporal pieces
# ------------------------------------------------------------------------------------------------
class TimeDistributed:
    """layers.py:208-264, style 'reshape1': (B,T,...) -> (B*T,...) -> model -> (B,T,...)."""

    def __init__(self, model, style="reshape1"):
        assert style in ["reshape1", "reshape2", "for"]
        self._style = style
        self.model = model

    def __call__(self, x, *args, **kwargs):
        B, T = x.shape[0], x.shape[1]
        y = self.model(x.reshape((B * T,) + tuple(x.shape[2:])), *args, **kwargs)
        if isinstance(y, (tuple, list)):
            return type(y)(yi.reshape((B, T) + tuple(yi.shape[1:])) if yi.shape[0] == B * T else yi for yi in y)
        return y.reshape((B, T) + tuple(y.shape[1:]))


class TemporalPooling:
    """layers.py:161-205, style 'direct': max / mean over axis 1 of (B,K,C,H,W) (bf16 carriers)."""

    def __init__(self, k, type="max", pool_size=None, strides=None, padding=0, style="direct"):
        assert type in ["max", "mean"]
        assert style in ["direct", "layer"]
        if pool_size is not None or style == "layer":
            raise NotImplementedError("only the 'direct' style (pool over the whole window) is on the hot path")
        self._type = type
        self._k = k

    def __call__(self, x):
        _require_cuda(x, "x")
        B, K = x.shape[0], x.shape[1]
        flat = x.reshape((B * K,) + tuple(x.shape[2:]))
        xb = to_nhwc_bf16(flat)                       # (B*K, C, H, W) channels-last
        _, C, H, W = xb.shape
        out = torch.empty((B, C, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        mode = _lib.VD_JOIN_MAX if self._type == "max" else _lib.VD_JOIN_MEAN
        check(load().vd_temporal_pool(ptr(xb), ptr(out), B, K, C * H * W, mode, stream_ptr()))
        return out


class TemporalTipConv:
    """The temporal cell of Conv('21', C, 3, 1, 1) (layers.py:82-89 second `_conv3d`):
    Conv3D(C, (3,1,1), pad (1,0,0), no bias) + BatchNorm(eps=1e-5) + LeakyReLU(0.1) on (B,T,C,H,W)."""

    def __init__(self, channels, epsilon=1e-5, slope=0.1, precision="bf16"):
        self.channels = channels
        self.epsilon = epsilon
        self.slope = slope
        self._precision = _precision_code(precision)   # fp32-parity modes: split carriers in and out (vd_temporal_conv_ex)
        self._w_split = {}
        self.weight = None                 # (Cout, Cin, 3, 1, 1) fp32, Gluon Conv3D layout
        self.gamma = torch.ones(channels, device="cuda")
        self.beta = torch.zeros(channels, device="cuda")
        self.running_mean = torch.zeros(channels, device="cuda")
        self.running_var = torch.ones(channels, device="cuda")
        self._w_taps = self._scale = self._shift = None

    def initialize(self, scale=0.07, generator=None):
        c = self.channels
        self.set_data((torch.rand((c, c, 3, 1, 1), generator=generator) * 2 - 1) * scale)
        return self

    def set_data(self, weight, gamma=None, beta=None, running_mean=None, running_var=None):
        c = self.channels
        self.weight = torch.as_tensor(weight).detach().to(torch.float32).cuda().reshape(c, c, 3)
        for name, v in (("gamma", gamma), ("beta", beta), ("running_mean", running_mean), ("running_var", running_var)):
            if v is not None:
                setattr(self, name, torch.as_tensor(v).detach().to(torch.float32).cuda().contiguous())
        # [tap][cout][cin] bf16 + folded inference BatchNorm
        self._w_taps = self.weight.permute(2, 0, 1).contiguous().to(torch.bfloat16)
        self._w_split = {}
        self._scale = (self.gamma / torch.sqrt(self.running_var + self.epsilon)).contiguous()
        self._shift = (self.beta - self.running_mean * self._scale).contiguous()

    def weight_split(self, planes):
        """(planes, 3, Cout, Cin) bf16 planes of the fp32 weight, [plane][tap][cout][cin] (vd_temporal_conv_ex)."""
        if planes not in self._w_split:
            c = self.channels
            taps = self.weight.permute(2, 0, 1).contiguous().reshape(3 * c, c)          # fp32 [tap][cout][cin]
            rows = torch.empty((3 * c, planes, c), dtype=torch.bfloat16, device=taps.device)
            check(load().vd_split_f32_rows(ptr(taps), ptr(rows), 3 * c, c, planes, stream_ptr()))
            self._w_split[planes] = rows.permute(1, 0, 2).contiguous().reshape(planes, 3, c, c)
        return self._w_split[planes]

    def split_forward(self, xs, B, T):
        """fp32-parity mode: xs = SplitF32 of the (B*T, C, H, W) window frames -> SplitF32 of the cell's fp32 output."""
        planes = _lib.PLANES[self._precision]
        assert planes > 1 and xs.planes == planes
        F, C, H, W = xs.shape
        assert F == B * T and C == self.channels
        y = torch.empty_like(xs.data)
        check(load().vd_temporal_conv_ex(ptr(xs.data), ptr(y), B, T, H, W, C, ptr(self.weight_split(planes)), ptr(self._scale),
                                         ptr(self._shift), float(self.slope), self._precision, 0, stream_ptr()))
        return SplitF32(y, xs.shape)

    def __call__(self, x):
        """x (B,T,C,H,W) -> (B,T,C,H,W) bf16 (channels-last per frame); in an fp32-parity mode: fp32 (B,T,C,H,W) in, SplitF32 of
        the (B*T,C,H,W) output frames out."""
        _require_cuda(x, "x")
        B, T, C, H, W = x.shape
        if self._precision != _lib.VD_PREC_BF16:
            return self.split_forward(to_nhwc_split(x.reshape(B * T, C, H, W), _lib.PLANES[self._precision]), B, T)
        xb = to_nhwc_bf16(x.reshape(B * T, C, H, W))
        y = torch.empty((B * T, C, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        check(load().vd_temporal_conv(ptr(xb), ptr(y), B, T, H, W, C, ptr(self._w_taps), ptr(self._scale),
                                      ptr(self._shift), float(self.slope), stream_ptr()))
        return y.reshape(B, T, C, H, W)


# ------------------------------------------------------------------------------------------------
# detection block (the convs in front of the tip)
# ------------------------------------------------------------------------------------------------
class ConvBNLReLU:
    """`_conv2d` (layers.py:63-70) / `_conv3d` (layers.py:73-79): Conv(no bias, stride 1, zero 'same' padding) +
    BatchNorm(eps 1e-5, inference statistics) + LeakyReLU(0.1).  kernel = k or (kh,kw) or (kt,kh,kw), extents 1 or 3.
    Parameters in Gluon layout: weight (Cout,Cin,kh,kw) / (Cout,Cin,kt,kh,kw), gamma, beta, running_mean, running_var."""

    def __init__(self, in_channels, channel, kernel, epsilon=1e-5, slope=0.1):
        if isinstance(kernel, int):
            kernel = (1, kernel, kernel)
        kernel = tuple(kernel)
        if len(kernel) == 2:
            kernel = (1,) + kernel
        assert len(kernel) == 3 and all(k in (1, 3) for k in kernel), "kernel extents must be 1 or 3"
        self.in_channels, self.channels, self.kernel = in_channels, channel, kernel
        self.epsilon, self.slope = epsilon, slope
        self.weight = None
        self.gamma = torch.ones(channel, device="cuda")
        self.beta = torch.zeros(channel, device="cuda")
        self.running_mean = torch.zeros(channel, device="cuda")
        self.running_var = torch.ones(channel, device="cuda")
        self._w_taps = self._scale = self._shift = None

    def initialize(self, scale=0.07, generator=None):
        shape = (self.channels, self.in_channels) + self.kernel
        self.set_data((torch.rand(shape, generator=generator) * 2 - 1) * scale)
        return self

    def set_data(self, weight, gamma=None, beta=None, running_mean=None, running_var=None):
        co, ci = self.channels, self.in_channels
        kt, kh, kw = self.kernel
        self.weight = torch.as_tensor(weight).detach().to(torch.float32).cuda().reshape(co, ci, kt, kh, kw)
        for name, v in (("gamma", gamma), ("beta", beta), ("running_mean", running_mean), ("running_var", running_var)):
            if v is not None:
                setattr(self, name, torch.as_tensor(v).detach().to(torch.float32).cuda().contiguous())
        # [tap][cout][cin] bf16, tap = (it*kh + iy)*kw + ix; folded inference BatchNorm
        self._w_taps = self.weight.permute(2, 3, 4, 0, 1).reshape(kt * kh * kw, co, ci).contiguous().to(torch.bfloat16)
        self._scale = (self.gamma / torch.sqrt(self.running_var + self.epsilon)).contiguous()
        self._shift = (self.beta - self.running_mean * self._scale).contiguous()

    def __call__(self, x):
        """x (B,Cin,H,W) or (B,T,Cin,H,W) -> same leading shape with Cout channels, bf16, channels-last per frame."""
        _require_cuda(x, "x")
        five = x.dim() == 5
        if not five:
            assert self.kernel[0] == 1, "a kernel with a temporal extent needs (B,T,C,H,W) input"
        B, T = (x.shape[0], x.shape[1]) if five else (x.shape[0], 1)
        C, H, W = x.shape[-3:]
        assert C == self.in_channels, "expected %d input channels, got %d" % (self.in_channels, C)
        assert self._w_taps is not None, "call initialize() or set_data() first"
        xb = to_nhwc_bf16(x.reshape(B * T, C, H, W))
        y = torch.empty((B * T, self.channels, H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        kt, kh, kw = self.kernel
        check(load().vd_conv_bn_lrelu(ptr(xb), ptr(y), B, T, H, W, C, self.channels, kt, kh, kw, ptr(self._w_taps),
                                      ptr(self._scale), ptr(self._shift), float(self.slope), stream_ptr()))
        return y.reshape(B, T, self.channels, H, W) if five else y


class _Seq:
    def __init__(self, cells):
        self.cells = list(cells)

    def __call__(self, x):
        for c in self.cells:
            x = c(x)
        return x

    def __iter__(self):
        return iter(self.cells)

    def __len__(self):
        return len(self.cells)

    def __getitem__(self, i):
        return self.cells[i]


def _expand_cell(conv_type, in_channels, channel):
    """The 3x3 'expand' of the block: `_conv2d(channel,3,1,1)` / `Conv('3',...)` = 3x3x3 / `Conv('21',...)` =
    (1,3,3) in->channel then (3,1,1) channel->channel (`_conv21d` with m=channel, layers.py:82-89,154-156)."""
    if conv_type == "2":
        return ConvBNLReLU(in_channels, channel, (1, 3, 3))
    if conv_type == "3":
        return ConvBNLReLU(in_channels, channel, (3, 3, 3))
    return _Seq([ConvBNLReLU(in_channels, channel, (1, 3, 3)), ConvBNLReLU(channel, channel, (3, 1, 1))])


class YOLODetectionBlockV3:
    """yolo3_temporal.py:184-239 (2-D twin in yolo3.py): body = [1x1 reduce -> channel, 3x3 expand -> 2*channel] x 2 +
    1x1 reduce; tip = 3x3 expand.  `__call__(x)` returns (route, tip).  conv_type '2': x (B,Cin,H,W); '3' / '21':
    x (B,T,Cin,H,W) (the reference swaps to (B,C,T,H,W) around the convs and back, :231-239 -- a no-op on the
    channels-last carrier).  Gluon infers `in_channels` at the first call; here it is a constructor argument."""

    def __init__(self, channel, conv_type="2", in_channels=None, **kwargs):
        assert conv_type in ["2", "3", "21"]
        assert channel % 2 == 0, "channel {} cannot be divided by 2".format(channel)
        self._conv_type = conv_type
        self.channel = channel
        cin = in_channels if in_channels is not None else channel * 2
        cells = []
        for _ in range(2):
            cells.append(ConvBNLReLU(cin, channel, (1, 1, 1)))
            cells.append(_expand_cell(conv_type, channel, channel * 2))
            cin = channel * 2
        cells.append(ConvBNLReLU(cin, channel, (1, 1, 1)))
        self.body = _Seq(cells)
        self.tip = _expand_cell(conv_type, channel, channel * 2)

    def cells(self):
        """All conv-BN-LReLU cells in execution order (body then tip), composite (2+1)D cells flattened."""
        out = []
        for c in list(self.body) + [self.tip]:
            out.extend(c.cells if isinstance(c, _Seq) else [c])
        return out

    def initialize(self, scale=0.07, generator=None):
        for c in self.cells():
            c.initialize(scale, generator)
        return self

    def __call__(self, x):
        _require_cuda(x, "x")
        assert x.dim() == (4 if self._conv_type == "2" else 5)
        route = self.body(x)
        tip = self.tip(route)
        return route, tip

    hybrid_forward = __call__


# ------------------------------------------------------------------------------------------------
# detector tail
# ------------------------------------------------------------------------------------------------
class YOLOV3Head:
    """Inference tail of YOLOV3 / YOLOV3T / YOLOV3Temporal: three YOLOOutputV3 blocks, scale concat,
    box_nms, post_nms slice, id/score/bbox split (yolo3.py:496,522-534; yolo3_temporal.py:468,542-555).

    `__call__(tips)` with tips = [s32, s16, s8] feature maps (B,Cin,H,W) -- or (B,T,Cin,H,W) for
    TimeDistributed heads -- returns ids (B[,T],post_nms,1), scores (B[,T],post_nms,1),
    bboxes (B[,T],post_nms,4).  One fused C-ABI call (vd_head_forward).

    temporal: None | 'conv21' (TemporalTipConv per scale in front of the output blocks, cfg T1)
              | 'cat' | 'max' | 'mean' (late joins of yolo3.py:1134-1138 collapsing K frames).
    """

    def __init__(self, classes, anchors=None, strides=None, channels=None, nms_thresh=0.45, nms_topk=400,
                 post_nms=100, temporal=None, k=1, agnostic=False, precision="bf16", fuse_tip=True, pair_kernel=True):
        """precision 'bf16' (default): bf16 operands, fp32 accumulate -- 1e-3 relative vs the fp32 reference on bf16-representable
        inputs.  precision 'fp32': fp32 NCHW tips / weights are split into three bf16 planes, six plane products -- 1e-5
        relative (VD_PREC_FP32_SPLIT); 'bf16x2': two planes, three products -- 1e-4 (VD_PREC_BF16X2).  Per-frame heads only."""
        self._precision = _precision_code(precision)
        self.pair_kernel = bool(pair_kernel)   # wide heads (31..80 classes): speculative head kernel on CTA pairs (bit-identical results; False = 1-CTA kernel)
        self.fuse_tip = bool(fuse_tip)     # temporal='conv21': tip cell + head in one kernel per scale where it applies (bit-identical; False = separate kernels)
        if self._precision != _lib.VD_PREC_BF16 and temporal not in (None, "conv21"):
            raise NotImplementedError("the fp32-parity modes cover the per-frame head and the 'conv21' temporal head")
        self.classes = list(classes) if not isinstance(classes, int) else list(range(classes))
        self._num_class = len(self.classes)
        anchors = DEFAULT_ANCHORS if anchors is None else anchors
        strides = DEFAULT_STRIDES if strides is None else strides
        # ORDER CONTRACT: anchors / strides / channels are given in OUTPUT order (deep -> shallow: s32, s16, s8), i.e. what
        # YOLOV3.__init__ holds AFTER its `anchors[::-1]`, `strides[::-1]` (yolo3.py:416-417).  wrappers.py:80-84 lists them
        # shallow -> deep (s8, s16, s32): reverse such lists before passing them here -- checked, not silently mis-paired.
        if list(strides) != sorted(strides, reverse=True):
            raise ValueError("YOLOV3Head: strides must be in output order (deep -> shallow, e.g. [32, 16, 8]); the reference's "
                             "constructor lists (s8, s16, s32) are reversed at yolo3.py:416-417 -- pass anchors[::-1], strides[::-1]")
        assert len(anchors) == len(strides), "one anchor list per stride"
        self.channels = DEFAULT_CHANNELS if channels is None else list(channels)
        assert temporal in (None, "conv21", "cat", "max", "mean")
        self.temporal, self.k = temporal, k
        self.nms_thresh, self.nms_topk, self.post_nms = nms_thresh, nms_topk, post_nms
        self.valid_thresh = 0.01                      # hard-coded at yolo3.py:527
        if agnostic:
            raise NotImplementedError("agnostic detector tail: use YOLOOutputV3(agnostic=True) + box_nms")
        self.yolo_outputs = [YOLOOutputV3(i, self._num_class, a, s, precision=precision) for i, (a, s) in enumerate(zip(anchors, strides))]
        self.tip_convs = [TemporalTipConv(c, precision=precision) for c in self.channels] if temporal == "conv21" else None
        self.pool = TemporalPooling(k, temporal) if temporal in ("max", "mean") else None
        self._keep = None
        self._ws_cache = {}

    def initialize(self, generator=None):
        mult = self.k if self.temporal == "cat" else 1
        for o, c in zip(self.yolo_outputs, self.channels):
            o.initialize(c * mult, generator=generator)
        if self.tip_convs:
            for t in self.tip_convs:
                t.initialize(generator=generator)
        return self

    def set_nms(self, nms_thresh=0.45, nms_topk=400, post_nms=100):
        """yolo3.py:536-556."""
        self.nms_thresh, self.nms_topk, self.post_nms = nms_thresh, nms_topk, post_nms

    def reset_class(self, classes, reuse_weights=None):
        self.classes = list(classes)
        self._num_class = len(self.classes)
        for o in self.yolo_outputs:
            o.reset_class(classes, reuse_weights)

    # -- helpers ---------------------------------------------------------------------------------
    def _ws(self, nbytes, device, tag):
        """The fused call keeps state in its workspace between calls (include/viddet_b200.h "Workspace contract"): one
        buffer per (head instance, device, stream, entry point), owned by the instance (freed with it)."""
        key = (device.index if device.index is not None else torch.cuda.current_device(), tag,
               torch.cuda.current_stream(device).cuda_stream)
        buf = self._ws_cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            self._ws_cache[key] = buf
        return buf

    def _params(self, tips, scratch):
        """Build VdHeadParams for already-channels-last tips [(frames,C,H,W) or (B,K,C,H,W)-flattened]."""
        p = VdHeadParams()
        p.num_scales = len(tips)
        p.num_class = self._num_class
        p.T = 1
        p.K_frames, p.join = 1, _lib.VD_JOIN_NONE
        p.nms_thresh, p.valid_thresh = float(self.nms_thresh), float(self.valid_thresh)
        p.nms_topk, p.post_nms = int(self.nms_topk), int(self.post_nms)
        p.precision = self._precision
        p.flags = (0 if self.fuse_tip else _lib.VD_HEAD_NO_FUSED_TIP) | (0 if self.pair_kernel else _lib.VD_HEAD_NO_PAIR_KERNEL)
        planes = _lib.PLANES[self._precision]
        split = planes > 1
        frames = None
        for i, (t, o) in enumerate(zip(tips, self.yolo_outputs)):
            F, C, H, W = t.shape
            s = p.scale[i]
            if self.temporal == "cat":
                assert F % self.k == 0
                p.K_frames, p.join = self.k, _lib.VD_JOIN_CAT
                F = F // self.k
            frames = F if frames is None else frames
            assert frames == F, "all scales must carry the same number of frames"
            s.tip_nhwc_bf16 = t.data.data_ptr() if split else t.data_ptr()
            s.tip_window_stride_frames = 1 if isinstance(t, ClipWindows) else 0
            s.weight_bf16 = (o.prediction.weight_split(planes) if split else o.prediction.weight_bf16).data_ptr()
            s.bias = o.prediction.bias.data_ptr()
            s.H, s.W, s.Cin = H, W, C
            s.stride = float(o._stride)
            for j in range(6):
                s.anchors[j] = float(o._anchors_np[j])
            if self.tip_convs is not None and not split:
                tc = self.tip_convs[i]
                s.tconv_weight_bf16 = tc._w_taps.data_ptr()
                s.tconv_scale = tc._scale.data_ptr()
                s.tconv_shift = tc._shift.data_ptr()
                s.tconv_out_nhwc_bf16 = scratch[i].data_ptr()
        p.frames = frames
        return p

    def _prepare(self, tips):
        lead = None
        flat = []
        planes = _lib.PLANES[self._precision]
        split = planes > 1
        for t in tips:
            if isinstance(t, ClipWindows):
                if self.temporal != "conv21" or split:
                    raise _lib.VidDetError(-1, "ClipWindows feed the bf16 temporal ('conv21') head; use .materialise() elsewhere")
                lead = (t.count, t.T)
                flat.append(t)
                continue
            if isinstance(t, SplitF32):
                assert split and t.planes == planes, "SplitF32 tips need a matching fp32-parity precision"
                flat.append(t)
                continue
            _require_cuda(t, "tip")
            if t.dim() == 5:
                lead = (t.shape[0], t.shape[1])
                t = t.reshape((t.shape[0] * t.shape[1],) + tuple(t.shape[2:]))
            flat.append(to_nhwc_split(t, planes) if split else to_nhwc_bf16(t))
        T = 1
        if self.temporal in ("max", "mean"):
            assert lead is not None, "temporal pooling needs (B,K,C,H,W) tips"
            flat = [self.pool(f.reshape(lead + tuple(f.shape[1:]))) for f in flat]
            lead = None
        elif self.temporal == "cat":
            assert lead is not None and lead[1] == self.k
            lead = None
        elif self.temporal == "conv21":
            assert lead is not None, "the temporal tip cell needs (B,T,C,H,W) tips"
            T = lead[1]
            if split:      # parity modes: the tip cells run here (split carriers in and out), the fused call then sees a plain per-frame head
                flat = [tc.split_forward(f, lead[0], T) for tc, f in zip(self.tip_convs, flat)]
                T = 1
        scratch = None
        if self.tip_convs is not None and not split:
            scratch = [torch.empty(tuple(f.shape), dtype=torch.bfloat16, device=f.device, memory_format=torch.channels_last) for f in flat]
        p = self._params(flat, scratch)
        p.T = T
        return p, flat, scratch, lead

    def __call__(self, tips, return_keep=False):
        if not (0 < self.nms_thresh < 1) or self.post_nms <= 0:
            # yolo3.py:525/529: NMS (or the slice) disabled -> plain rows
            det = self.detections(tips)
            if 0 < self.nms_thresh < 1:
                det = box_nms(det, overlap_thresh=self.nms_thresh, valid_thresh=self.valid_thresh, topk=self.nms_topk,
                              id_index=0, score_index=1, coord_start=2, force_suppress=False)
            return det[..., 0:1], det[..., 1:2], det[..., 2:]
        p, flat, scratch, lead = self._prepare(tips)
        dev = flat[0].device
        F, post = p.frames, self.post_nms
        ids = torch.empty((F, post, 1), device=dev)
        scores = torch.empty((F, post, 1), device=dev)
        bboxes = torch.empty((F, post, 4), device=dev)
        keep = torch.empty((F, post), dtype=torch.int32, device=dev) if return_keep else None
        lib = load()
        ws = self._ws(lib.vd_head_workspace_bytes(ctypes.byref(p)), dev, "head")
        check(lib.vd_head_forward(ctypes.byref(p), ptr(ids), ptr(scores), ptr(bboxes), ptr(keep), ptr(ws), ws.numel(),
                                  stream_ptr()))
        if lead is not None:
            ids, scores, bboxes = (t.reshape(lead + tuple(t.shape[1:])) for t in (ids, scores, bboxes))
            if keep is not None:
                keep = keep.reshape(lead + (post,))
        return (ids, scores, bboxes, keep) if return_keep else (ids, scores, bboxes)

    def session(self, tips, return_keep=False, out=None, mirrors=None):
        """Bind tips/outputs/workspace once; the returned HeadSession re-enqueues the same call with no
        per-call Python work (and can be captured into a CUDA graph).  mirrors: byte deltas to the same output slot in the
        peer GPUs' gather buffers (viddet_b200.dist.PeerGather.deltas) -- the NMS kernel then stores every result there too."""
        return HeadSession(self, tips, return_keep, out, mirrors)

    def train_outputs(self, tips):
        """The training-mode branch of YOLOV3.hybrid_forward without a recorded loss (yolo3.py:498-509,532-535): per scale the
        output layer's 7-tuple, the raw parts flattened with reshape((0,-3,-1)) and concatenated over the scales (deep -> shallow).
        Returns (box_preds (B,N,4), all_anchors, all_offsets, all_feat_maps [(1,1,H,W) fake maps], box_centers (B,N,2),
        box_scales (B,N,2), objness (B,N,1), class_pred (B,N,C)), N = 3*sum(HW) -- the row order of the prefetched targets."""
        dets, centers, scales, objs, clss, ancs, offs, fmaps = [], [], [], [], [], [], [], []
        for t, o in zip(tips, self.yolo_outputs):
            _require_cuda(t, "tip")
            bbox, rc, rs, ob, cp, an, of = o(t, training=True)
            B = bbox.shape[0]
            dets.append(bbox); centers.append(rc.reshape(B, -1, 2)); scales.append(rs.reshape(B, -1, 2))
            objs.append(ob.reshape(B, -1, 1)); clss.append(cp.reshape(B, -1, cp.shape[-1]))
            ancs.append(an); offs.append(of)
            fmaps.append(torch.zeros((1, 1, t.shape[-2], t.shape[-1]), device=bbox.device))
        cat = lambda xs: torch.cat(xs, dim=1)
        return cat(dets), ancs, offs, fmaps, cat(centers), cat(scales), cat(objs), cat(clss)

    def train_forward(self, tips, gt_boxes, obj_t, centers_t, scales_t, weights_t, clas_t, ignore_iou_thresh=0.7):
        """The recorded training branch (yolo3.py:510-516): losses = YOLOV3Loss(preds + YOLOV3TargetMerger(box_preds, gt, prefetched
        targets)) -> (obj_loss, center_loss, scale_loss, cls_loss), forward values, all on the device."""
        box_preds, _, _, _, centers, scales, objness, cls_pred = self.train_outputs(tips)
        merged = YOLOV3TargetMerger(self._num_class, ignore_iou_thresh)(box_preds, gt_boxes, obj_t, centers_t, scales_t, weights_t, clas_t)
        return YOLOV3Loss()(objness, centers, scales, cls_pred, *merged)

    def detections(self, tips):
        """The concatenated (B[,T], rows, 6) tensor of yolo3.py:523 (before NMS)."""
        p, flat, scratch, lead = self._prepare(tips)
        dev = flat[0].device
        rows = sum(self._num_class * int(p.scale[i].H) * int(p.scale[i].W) * 3 for i in range(p.num_scales))
        det = torch.empty((p.frames, rows, 6), device=dev)
        lib = load()
        ws = self._ws(max(lib.vd_head_workspace_bytes(ctypes.byref(p)), 256), dev, "det")
        check(lib.vd_head_detections(ctypes.byref(p), ptr(det), ptr(ws), ws.numel(), stream_ptr()))
        if lead is not None:
            det = det.reshape(lead + (rows, 6))
        return det


class HeadSession:
    """A YOLOV3Head call with everything pre-bound: static channels-last bf16 input buffers
    (`self.tips`, refill with copy_), static outputs (`ids`, `scores`, `bboxes`, `keep`), private
    workspace.  `run()` enqueues the kernels on the current stream; `capture()` records them into a
    CUDA graph whose `replay()` is a single launch."""

    def __init__(self, head, tips, return_keep=False, out=None, mirrors=None):
        if not (0 < head.nms_thresh < 1) or head.post_nms <= 0:
            raise _lib.VidDetError(-1, "HeadSession needs NMS and post_nms enabled")
        self.head = head
        self.params, self.tips, self._scratch, self.lead = head._prepare(tips)
        if mirrors:
            assert out is not None, "output mirrors need caller-owned outputs inside the gather buffer"
            assert len(mirrors) <= _lib.VD_MAX_MIRRORS
            self.params.n_mirrors = len(mirrors)
            for i, d in enumerate(mirrors):
                self.params.mirror_delta[i] = int(d)
        dev = self.tips[0].device
        F, post = self.params.frames, head.post_nms
        if out is not None:                           # caller-owned output buffers (e.g. slices of one tensor per ring)
            self.ids, self.scores, self.bboxes = out
            assert tuple(self.ids.shape) == (F, post, 1) and tuple(self.scores.shape) == (F, post, 1) and tuple(self.bboxes.shape) == (F, post, 4)
            assert self.ids.is_contiguous() and self.scores.is_contiguous() and self.bboxes.is_contiguous()
        else:
            self.ids = torch.empty((F, post, 1), device=dev)
            self.scores = torch.empty((F, post, 1), device=dev)
            self.bboxes = torch.empty((F, post, 4), device=dev)
        self.keep = torch.empty((F, post), dtype=torch.int32, device=dev) if return_keep else None
        lib = load()
        self._ws = torch.zeros(lib.vd_head_workspace_bytes(ctypes.byref(self.params)) + 256, dtype=torch.uint8, device=dev)
        self.launches = lib.vd_head_launch_count(ctypes.byref(self.params))
        self.fused_tip = lib.vd_head_fused_tip(ctypes.byref(self.params)) == 1      # temporal tip cell + head as one kernel per scale
        self.graph = None
        self._graphs = {}

    def run(self, stage_mask=_lib.VD_STAGE_ALL):
        check(load().vd_head_forward_stages(ctypes.byref(self.params), ptr(self.ids), ptr(self.scores), ptr(self.bboxes),
                                            ptr(self.keep), ptr(self._ws), self._ws.numel(), stream_ptr(), int(stage_mask)))
        return self.ids, self.scores, self.bboxes

    def capture(self, stage_mask=_lib.VD_STAGE_ALL):
        """Record the call (or the stages in `stage_mask`) into a CUDA graph; `replay(stage_mask)` launches it."""
        self.run(stage_mask)                          # warm-up outside capture (function attributes, lazy init)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(stage_mask)
        self._graphs[int(stage_mask)] = g
        if stage_mask == _lib.VD_STAGE_ALL:
            self.graph = g
        return self

    def replay(self, stage_mask=_lib.VD_STAGE_ALL):
        self._graphs[int(stage_mask)].replay()
        return self.ids, self.scores, self.bboxes

    def rebind(self, tips):
        """Point the session at other resident input tensors of the same shapes (already channels-last bf16); the workspace
        (thresholds, hints) is kept.  Not for captured graphs: they hold the old pointers."""
        assert len(tips) == len(self.tips)
        for i, (t, old) in enumerate(zip(tips, self.tips)):
            assert tuple(t.shape) == tuple(old.shape) and type(t) is type(old)
            if not isinstance(t, ClipWindows):
                assert t.dtype == torch.bfloat16 and t.is_contiguous(memory_format=torch.channels_last)
            self.params.scale[i].tip_nhwc_bf16 = t.data_ptr()
        self.tips = list(tips)
        return self

    def stats(self):
        """(frames redone by the exact path, completed calls) -- running totals of this workspace (synchronises)."""
        off = load().vd_head_stats_offset(ctypes.byref(self.params))
        torch.cuda.synchronize()
        w = self._ws[off + 20: off + 28].view(torch.int32).tolist()
        return int(w[0]) & 0xffffffff, int(w[1]) & 0xffffffff

    def redone_frames(self):
        """Frames of the last completed call that the exact path had to redo (synchronises; 0 in the steady state)."""
        off = load().vd_head_stats_offset(ctypes.byref(self.params))
        torch.cuda.synchronize()
        return int(self._ws[off + 16: off + 20].view(torch.int32).item())

    def packed(self):
        """(frames, post_nms, 6) rows [id, score, x1, y1, x2, y2] -- the box_nms row layout."""
        return torch.cat([self.ids, self.scores, self.bboxes], dim=-1)


class HeadPipeline:
    """Throughput mode over a ring of HeadSessions (one per in-flight batch).

    One CUDA graph holds `rotations` rotations of the ring.  The fused head kernels (pred conv + decode +
    candidate filter) of consecutive batches run back to back on one stream; the per-frame top-k/NMS kernel of
    batch j runs on a second stream as soon as its head kernel has finished, i.e. concurrently with the head
    kernel of batch j+1, whose SMs it shares (it is sized for that: 256 threads, 56 registers, ~30 KB shared
    memory; the head kernel's dynamic tile scheduler absorbs the interference).  A session's next head kernel
    waits for its previous NMS kernel.  `cycle()` = rotations * len(sessions) steps; every batch is complete
    when the graph has finished."""

    def __init__(self, sessions, rotations=2, steps=None, inputs=None):
        """steps: batches per graph (default rotations * len(sessions); any count >= 1: step i runs on session i % len).
        inputs: optional list of resident input sets (each a list of channels-last bf16 tips of the sessions' shapes); step i
        then reads inputs[i % len(inputs)] while using session i % len(sessions)'s workspace and outputs -- every threshold the
        speculative filter uses was learned from a DIFFERENT batch than the one it filters (fresh data every step)."""
        assert len(sessions) >= 2 and rotations >= 1
        self.sessions = list(sessions)
        steps = rotations * len(self.sessions) if steps is None else int(steps)
        assert steps >= 1
        self._streams = [torch.cuda.Stream() for _ in range(2)]
        for s in self.sessions:                       # initialise workspaces (scheduler state, warm-start hints)
            s.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream()
            hs, ns = self._streams
            hs.wait_stream(main)
            ns.wait_stream(main)
            nms_done = [None] * len(self.sessions)
            for i in range(steps):
                j = i % len(self.sessions)
                sess = self.sessions[j]
                if inputs is not None:
                    sess.rebind(inputs[i % len(inputs)])            # the captured nodes keep this step's pointers
                with torch.cuda.stream(hs):
                    if nms_done[j] is not None:
                        hs.wait_event(nms_done[j])            # the session's buffers are free again
                    sess.run(_lib.VD_STAGE_TCONV | _lib.VD_STAGE_HEAD)   # (temporal tip cell kernels, if any,) fused head kernel
                    head_done = torch.cuda.Event()
                    head_done.record(hs)
                with torch.cuda.stream(ns):
                    ns.wait_event(head_done)
                    sess.run(_lib.VD_STAGE_NMS)
                    nms_done[j] = torch.cuda.Event()
                    nms_done[j].record(ns)
            main.wait_stream(hs)
            main.wait_stream(ns)
        self._graph = g
        self.steps_per_cycle = steps
        self.launches_per_step = 2

    def cycle(self):
        """Run rotations * len(sessions) batches, head and NMS kernels overlapped across batches."""
        self._graph.replay()


# ------------------------------------------------------------------------------------------------
# training targets
# ------------------------------------------------------------------------------------------------
class YOLOV3PrefetchTargetGenerator:
    """yolo_target.py:13-148.  `__call__(img, xs, anchors, offsets, gt_boxes, gt_ids, gt_mixratio=None)`
    -> objectness (B,N,1), center_targets (B,N,2), scale_targets (B,N,2), weights (B,N,2),
    class_targets (B,N,C), N = 3*sum(H_i*W_i), rows in the order of the train-mode predictions.

    img / xs are used for their shapes only (yolo_target.py:72-73,110-111) and may be tensors or
    shape tuples; anchors: 3 x (1,1,3,2); offsets: 3 x (1,HW_i,1,2) (sizes only)."""

    def __init__(self, num_class, **kwargs):
        self._num_class = num_class

    def alloc_outputs(self, B, N, device):
        """The five target tensors (objectness, center, scale, weights, class) for `run_into` (static buffers, e.g. for CUDA graphs)."""
        return tuple(torch.empty((B, N, w), device=device) for w in (1, 2, 2, 2, self._num_class))

    def run_into(self, img, xs, anchors, offsets, gt_boxes, gt_ids, gt_mixratio, outs):
        """`__call__` writing into caller-owned outputs (from alloc_outputs); returns them."""
        return self.__call__(img, xs, anchors, offsets, gt_boxes, gt_ids, gt_mixratio, out=outs)

    def __call__(self, img, xs, anchors, offsets, gt_boxes, gt_ids, gt_mixratio=None, return_assign=False, out=None):
        assert isinstance(anchors, (list, tuple)) and isinstance(offsets, (list, tuple)) and isinstance(xs, (list, tuple))
        assert len(xs) == len(anchors) == len(offsets) == 3
        _require_cuda(gt_boxes, "gt_boxes")
        _require_cuda(gt_ids, "gt_ids")
        ishape = tuple(img.shape) if hasattr(img, "shape") else tuple(img)
        hw = []
        for x, o in zip(xs, offsets):
            xshape = tuple(x.shape) if hasattr(x, "shape") else tuple(x)
            hw += [int(xshape[2]), int(xshape[3])]
            n_off = int(np.prod(tuple(o.shape))) // 2
            assert n_off == xshape[2] * xshape[3], "offsets must cover the feature map"
        anc = np.concatenate([np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a,
                                         dtype=np.float32).reshape(-1, 2) for a in anchors], 0)
        assert anc.shape == (9, 2)
        gb = gt_boxes.to(torch.float32).contiguous()
        gi = gt_ids.to(torch.float32).contiguous()
        mix = gt_mixratio.to(torch.float32).contiguous() if gt_mixratio is not None else None
        B, M = gb.shape[0], gb.shape[1]
        C = self._num_class
        N = 3 * sum(hw[2 * i] * hw[2 * i + 1] for i in range(3))
        dev = gb.device
        if out is not None:
            obj, ctr, scl, wgt, cls = out
            for t, w in zip(out, (1, 2, 2, 2, C)):
                assert tuple(t.shape) == (B, N, w) and t.is_contiguous() and t.dtype == torch.float32 and t.is_cuda
        else:
            obj, ctr, scl, wgt, cls = self.alloc_outputs(B, N, dev)
        match = torch.empty((B, M), dtype=torch.int32, device=dev) if return_assign else None
        row = torch.empty((B, M), dtype=torch.int32, device=dev) if return_assign else None
        hw_c = (ctypes.c_int * 6)(*hw)
        anc_c = (ctypes.c_float * 18)(*anc.reshape(-1).tolist())
        check(load().vd_prefetch_targets(B, M, C, int(ishape[2]), int(ishape[3]), hw_c, anc_c, ptr(gb), ptr(gi),
                                         int(gi.shape[-1]), ptr(mix), ptr(obj), ptr(ctr), ptr(scl), ptr(wgt), ptr(cls),
                                         ptr(match), ptr(row), stream_ptr()))
        outs = (obj, ctr, scl, wgt, cls)
        return outs + (match, row) if return_assign else outs

    forward = __call__


# ------------------------------------------------------------------------------------------------
# training: dynamic targets, target merger, loss
# ------------------------------------------------------------------------------------------------
class YOLOV3TargetMerger:
    """yolo_target.py:208-281.  `__call__(box_preds, gt_boxes, obj_t, centers_t, scales_t, weights_t, clas_t)` ->
    [objectness, center_targets, scale_targets, weights, class_targets, class_mask].  One vd_target_merge call."""

    def __init__(self, num_class, ignore_iou_thresh, **kwargs):
        self._num_class = num_class
        self._ignore_iou_thresh = float(ignore_iou_thresh)
        self._label_smooth = False

    def _run(self, box_preds, gt_boxes, prefetched):
        _require_cuda(box_preds, "box_preds")
        _require_cuda(gt_boxes, "gt_boxes")
        bp = box_preds.to(torch.float32).reshape(box_preds.shape[0], -1, 4).contiguous()
        gt = gt_boxes.to(torch.float32).contiguous()
        B, N = bp.shape[0], bp.shape[1]
        M, C, dev = gt.shape[1], self._num_class, bp.device
        pre = [None] * 5
        if prefetched is not None:
            shapes = [(B, N, 1), (B, N, 2), (B, N, 2), (B, N, 2), (B, N, C)]
            pre = []
            for t, shp in zip(prefetched, shapes):
                _require_cuda(t, "prefetched target")
                assert tuple(t.shape) == shp, "prefetched target shape %s != %s" % (tuple(t.shape), shp)
                pre.append(t.to(torch.float32).contiguous())
        outs = [torch.empty((B, N, w), device=dev) for w in (1, 2, 2, 2, C, C)]
        check(load().vd_target_merge(B, N, M, C, ptr(bp), ptr(gt), *[ptr(t) for t in pre], self._ignore_iou_thresh,
                                     int(self._label_smooth), *[ptr(t) for t in outs], stream_ptr()))
        return outs

    def __call__(self, box_preds, gt_boxes, obj_t, centers_t, scales_t, weights_t, clas_t):
        return self._run(box_preds, gt_boxes, (obj_t, centers_t, scales_t, weights_t, clas_t))

    forward = __call__


class YOLOV3DynamicTargetGeneratorSimple(YOLOV3TargetMerger):
    """yolo_target.py:151-205: targets that depend on the current predictions only (pos_iou_thresh >= 1):
    objectness -1 where the best IoU with a ground truth exceeds ignore_iou_thresh, zeros / -1 elsewhere."""

    def __call__(self, box_preds, gt_boxes):
        return self._run(box_preds, gt_boxes, None)[:5]

    forward = __call__


class YOLOV3Loss:
    """gluoncv.loss.YOLOV3Loss forward (constructed at yolo3.py:409, called at :515):
    `__call__(objness, box_centers, box_scales, cls_preds, objness_t, center_t, scale_t, weight_t, class_t, class_mask)`
    -> (obj_loss, center_loss, scale_loss, cls_loss), each (B,)."""

    def __init__(self, batch_axis=0, weight=None, **kwargs):
        assert batch_axis == 0 and weight is None

    def __call__(self, objness, box_centers, box_scales, cls_preds, objness_t, center_t, scale_t, weight_t, class_t, class_mask):
        ts = [objness, box_centers, box_scales, cls_preds, objness_t, center_t, scale_t, weight_t, class_t, class_mask]
        for t in ts:
            _require_cuda(t, "loss input")
        ts = [t.to(torch.float32).contiguous() for t in ts]
        B, N, C = cls_preds.shape[0], cls_preds.shape[1], cls_preds.shape[2]
        for t, w in zip(ts, (1, 2, 2, C, 1, 2, 2, 2, C, C)):
            assert t.numel() == B * N * w, "loss input has %d elements, expected %d" % (t.numel(), B * N * w)
        dev = ts[0].device
        outs = [torch.empty((B,), device=dev) for _ in range(4)]
        lib = load()
        ws = _workspace(max(lib.vd_yolo3_loss_workspace_bytes(B, N), 256), dev, owner="loss")
        check(lib.vd_yolo3_loss(B, N, C, *[ptr(t) for t in ts], *[ptr(t) for t in outs], ptr(ws), ws.numel(), stream_ptr()))
        return tuple(outs)

    forward = __call__


# ------------------------------------------------------------------------------------------------
# post-processing of detect()
# ------------------------------------------------------------------------------------------------
def upsample_concat(x, route):
    """yolo3.py:515-519 / yolo3_temporal.py:500-506: concat(slice_like(_upsample(x, 2), route, axes=(-2,-1)), route, dim=-3) on
    channels-last bf16 carriers; (B,C,H,W) or (B,T,C,H,W) (frames are independent)."""
    _require_cuda(x, "x"); _require_cuda(route, "route")
    if x.dim() == 5:
        B_, T_ = x.shape[0], x.shape[1]
        out = upsample_concat(x.reshape((B_ * T_,) + tuple(x.shape[2:])), route.reshape((B_ * T_,) + tuple(route.shape[2:])))
        return out.reshape((B_, T_) + tuple(out.shape[1:]))
    xb, rb = to_nhwc_bf16(x), to_nhwc_bf16(route)
    B, C1, H, W = xb.shape
    B2, C2, H2, W2 = rb.shape
    assert B == B2, "batch mismatch"
    out = torch.empty((B, C1 + C2, H2, W2), dtype=torch.bfloat16, device=xb.device, memory_format=torch.channels_last)
    check(load().vd_upsample_concat(ptr(xb), ptr(rb), ptr(out), B, H, W, C1, H2, W2, C2, stream_ptr()))
    return out


class YOLOV3Neck:
    """Everything `YOLOV3.hybrid_forward` does after the backbone stages (yolo3.py:496-534, inference): per scale, from deep to
    shallow, YOLODetectionBlockV3 -> (route, tip); the tip goes to the output layer, the route through the transition
    `_conv2d(channel, 1, 0, 1)` (yolo3.py:425-427), x2 upsample and concat in front of the next backbone route; then the fused
    head tail (decode, concat, box_nms, slice).  `routes` = backbone stage outputs in the reference's stage order
    (shallow -> deep: s8, s16, s32), NCHW logical layout.  Same constructor vocabulary as YOLOV3 (classes, channels in output
    order, nms parameters); Gluon infers the input channels lazily, here `stage_channels` (shallow -> deep) states them."""

    def __init__(self, classes, channels=(512, 256, 128), stage_channels=(256, 512, 1024), anchors=None, strides=None,
                 nms_thresh=0.45, nms_topk=400, post_nms=100, conv_type="2"):
        """anchors / strides: OUTPUT order (deep -> shallow), i.e. the reference's lists after yolo3.py:416-417 reversed them
        (YOLOV3Head checks that strides descend); channels: as the reference passes them (deep -> shallow already, wrappers.py:91-107).
        conv_type '2': YOLOV3 (routes (B,C,H,W)); '3' / '21': YOLOV3Temporal with t_out (yolo3_temporal.py:448-555): routes
        (B,T,C,H,W), 3-D / (2+1)D detection blocks, transitions and output layers applied per frame (TimeDistributed), detections
        (B,T,post,.)."""
        assert conv_type in ("2", "3", "21")
        self._conv_type = conv_type
        n = len(channels)
        assert len(stage_channels) == n
        self.channels = list(channels)
        deep_first = list(stage_channels)[::-1]
        self.yolo_blocks, self.transitions = [], []
        for i, ch in enumerate(self.channels):
            cin = deep_first[i] if i == 0 else self.channels[i] + deep_first[i]      # transition i-1 emits channels[i] (yolo3.py:425-427)
            self.yolo_blocks.append(YOLODetectionBlockV3(ch, conv_type, in_channels=cin))
            if i > 0:
                self.transitions.append(ConvBNLReLU(self.channels[i - 1], ch, (1, 1, 1)))
        self.head = YOLOV3Head(classes, anchors=anchors, strides=strides, channels=[2 * c for c in self.channels],
                               nms_thresh=nms_thresh, nms_topk=nms_topk, post_nms=post_nms)
        self.yolo_outputs = self.head.yolo_outputs

    def initialize(self, scale=0.07, generator=None):
        for b in self.yolo_blocks:
            b.initialize(scale, generator)
        for t in self.transitions:
            t.initialize(scale, generator)
        self.head.initialize(generator=generator)
        return self

    def set_nms(self, nms_thresh=0.45, nms_topk=400, post_nms=100):
        self.head.set_nms(nms_thresh, nms_topk, post_nms)

    def tips(self, routes):
        """The three tips (deep -> shallow), channels-last bf16."""
        routes = list(routes)[::-1]                                   # yolo3.py:518 `routes[::-1]`
        x, tips = routes[0], []
        for i, block in enumerate(self.yolo_blocks):
            x, tip = block(x)
            tips.append(tip)
            if i >= len(routes) - 1:
                break
            if x.dim() == 5:                                           # TimeDistributed(self.transitions[i]) (yolo3_temporal.py:495-497)
                B_, T_ = x.shape[0], x.shape[1]
                x = self.transitions[i](x.reshape((B_ * T_,) + tuple(x.shape[2:])))
                x = x.reshape((B_, T_) + tuple(x.shape[1:]))
            else:
                x = self.transitions[i](x)
            x = upsample_concat(x, routes[i + 1])
        return tips

    def __call__(self, routes):
        return self.head(self.tips(routes))

    def detections(self, routes):
        """The (B, rows, 6) tensor `concat(all_detections)` holds at yolo3.py:523."""
        return self.head.detections(self.tips(routes))

    def session(self, routes):
        """Bind routes once and record the whole forward (19-28 conv cells, the glue kernels, the fused head) into ONE CUDA graph:
        `NeckSession.replay()` is a single launch with no per-call Python, allocation or tensor-map work."""
        return NeckSession(self, routes)


class NeckSession:
    """A YOLOV3Neck forward with static inputs (`routes`: refill with copy_), static outputs (`ids`, `scores`, `bboxes`) and
    every intermediate activation owned by the graph's memory pool.  The warm-up and the capture run on the same private stream,
    so the head's workspace (its thresholds live there between replays) is the one created during the warm-up, not a buffer the
    graph would re-zero."""

    def __init__(self, neck, routes):
        self.neck = neck
        self.routes = []
        for r in routes:
            _require_cuda(r, "route")
            if r.dim() == 4:
                self.routes.append(to_nhwc_bf16(r).clone(memory_format=torch.preserve_format))
            else:                                           # (B,T,C,H,W): channels-last per frame
                B, T = r.shape[0], r.shape[1]
                f = to_nhwc_bf16(r.reshape((B * T,) + tuple(r.shape[2:]))).clone(memory_format=torch.preserve_format)
                self.routes.append(f.reshape((B, T) + tuple(f.shape[1:])))
        self._stream = torch.cuda.Stream()
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(2):                              # function attributes, workspace, thresholds
                neck(self.routes)
        self._stream.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self._stream):
            self.ids, self.scores, self.bboxes = neck(self.routes)

    def replay(self):
        self.graph.replay()
        return self.ids, self.scores, self.bboxes


def postprocess_detections(ids, scores, bboxes, size):
    """detect_yolo3.py:222-261 on device: clip boxes to [0, size], keep rows with id >= 0 (order preserved), divide the
    boxes by size.  ids/scores (..., post, 1), bboxes (..., post, 4) -> rows (F, post, 6) [id, score, x1, y1, x2, y2]
    padded with -1 and counts (F,) int32, F = product of the leading dims (B, or B*T for mult_out windows)."""
    for t in (ids, scores, bboxes):
        _require_cuda(t, "detections")
    post = ids.shape[-2]
    i2 = ids.to(torch.float32).reshape(-1, post).contiguous()
    s2 = scores.to(torch.float32).reshape(-1, post).contiguous()
    b2 = bboxes.to(torch.float32).reshape(-1, post, 4).contiguous()
    F = i2.shape[0]
    rows = torch.empty((F, post, 6), device=i2.device)
    counts = torch.empty((F,), dtype=torch.int32, device=i2.device)
    check(load().vd_postprocess_detections(ptr(i2), ptr(s2), ptr(b2), F, post, float(size), ptr(rows), ptr(counts), stream_ptr()))
    return rows, counts


class ClassTree:
    """The three tables `hierarchical_nms` reads from the combined dataset (detect_yolo3.py:738-746), on the device:
    levels = dataset.get_levels() (combined.py:117-126), parent = class index of each class's parent (-1 under ROOT),
    branch[i][j] = dataset.on_branch(i, j) (combined.py:143-150).  `from_parents(wn_classes, parents)` builds them from the
    reference's own encoding (list of wnids + child->parent dict with 'ROOT')."""

    def __init__(self, levels, parent, branch):
        self.levels = torch.as_tensor(np.asarray(levels), dtype=torch.int32).cuda().contiguous()
        self.parent = torch.as_tensor(np.asarray(parent), dtype=torch.int32).cuda().contiguous()
        self.branch = torch.as_tensor(np.asarray(branch), dtype=torch.uint8).cuda().contiguous()
        self.num_class = int(self.levels.numel())
        assert self.parent.numel() == self.num_class and tuple(self.branch.shape) == (self.num_class, self.num_class)
        par = np.asarray(parent).astype(np.int64)
        if ((par >= self.num_class) | (par < -1)).any():            # the device walks parent[] without bounds checks
            raise ValueError("ClassTree: parent indices must be in [-1, num_class)")
        if (par == np.arange(self.num_class)).any():
            raise ValueError("ClassTree: a class cannot be its own parent")
        self.min_level = int(np.asarray(levels).min()) if self.num_class else 0

    @staticmethod
    def tables(wn_classes, parents):
        """(levels, parent, branch) numpy tables from the reference's encoding (host-side, no device needed)."""
        idx = {c: i for i, c in enumerate(wn_classes)}
        n = len(wn_classes)
        parent = np.array([idx[parents[c]] if parents[c] != "ROOT" else -1 for c in wn_classes], np.int32)
        levels = np.zeros(n, np.int32)
        anc = []
        for i in range(n):
            chain, p = [i], parent[i]
            while p >= 0:
                chain.append(int(p)); p = parent[p]
            levels[i] = len(chain)
            anc.append(set(chain))
        branch = np.zeros((n, n), np.uint8)
        for i in range(n):
            for j in range(n):
                child, par = max(i, j), min(i, j)            # combined.py:147-150: the smaller index is looked up in the larger one's lineage
                branch[i, j] = 1 if (i == j or par in anc[child]) else 0
        return levels, parent, branch

    @classmethod
    def from_parents(cls, wn_classes, parents):
        return cls(*cls.tables(wn_classes, parents))


def hierarchical_nms(rows, counts, tree, ov_thresh=0.5, conf_thresh=0.0, level_thresh=10, arith="float64"):
    """detect_yolo3.py:736-789 on the packed predictions `postprocess_detections` returns: rows (F, post, 6), counts (F,) ->
    (rows, counts) of the merged lists.  arith 'float64' = the reference on re-loaded predictions (Python floats),
    'legacy32' = its in-memory path under the NumPy of its era (np.float32 scalars)."""
    _require_cuda(rows, "rows"); _require_cuda(counts, "counts")
    assert rows.dim() == 3 and rows.shape[-1] == 6
    level_thresh = max(0, int(level_thresh))
    if tree.num_class and level_thresh < tree.min_level:
        raise ValueError("'ROOT' is not in list")       # what `cls_map.index(parents[...])` raises at :766 once a lift reaches ROOT
    F, post, _ = rows.shape
    r = rows.to(torch.float32).contiguous()
    c = counts.to(torch.int32).contiguous()
    out = torch.empty_like(r)
    ocnt = torch.empty_like(c)
    check(load().vd_hierarchical_nms(ptr(r), ptr(c), F, post, tree.num_class, ptr(tree.levels), ptr(tree.parent), ptr(tree.branch),
                                     float(ov_thresh), float(conf_thresh), level_thresh, {"float64": 0, "legacy32": 1}[arith],
                                     ptr(out), ptr(ocnt), stream_ptr()))
    return out, ocnt
