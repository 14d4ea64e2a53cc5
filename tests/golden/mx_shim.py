"""A numpy-backed stand-in for the handful of MXNet / Gluon / GluonCV names the reference's hot-path classes touch, so that
their OWN Python (cut out of /root/reference with `ast`, never copied) can be executed here to generate golden vectors
(tests/golden/make_golden_ref_exec.py).  TEST INFRASTRUCTURE ONLY.  The operators below restate the published MXNet semantics
(fp32 everywhere, reshape codes 0/-1/-2/-3, float 0/1 comparison results, first-max argmax returned as fp32, corner-format
box_iou); what the goldens pin is the reference's own logic on top of them: slicing, row order, the target loop, index math."""
import contextlib

import numpy as np

f32 = np.float32


def _unwrap(v):
    return v.a if isinstance(v, ND) else v


class ND:
    __array_priority__ = 100

    def __init__(self, a):
        self.a = np.asarray(a, dtype=f32) if not (isinstance(a, np.ndarray) and a.dtype == f32) else a

    # ---- introspection
    shape = property(lambda s: tuple(s.a.shape))
    size = property(lambda s: int(s.a.size))
    ndim = property(lambda s: s.a.ndim)

    def asnumpy(self):
        return np.array(self.a, dtype=f32)

    # ---- shape ops
    def reshape(self, *shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if not shape:
            shape = tuple(kw["shape"])
        src, out, i = list(self.a.shape), [], 0
        for code in shape:
            if code == 0:
                out.append(src[i]); i += 1
            elif code == -1:
                out.append(-1); i += 1
            elif code == -2:
                out.extend(src[i:]); i = len(src)
            elif code == -3:
                out.append(src[i] * src[i + 1]); i += 2
            else:
                out.append(int(code)); i += 1
        return ND(self.a.reshape(out))

    def transpose(self, axes=None):
        return ND(np.ascontiguousarray(self.a.transpose(axes)))

    def slice_axis(self, axis, begin, end):
        sl = [slice(None)] * self.a.ndim
        sl[axis] = slice(begin, end)
        return ND(np.array(self.a[tuple(sl)]))

    def expand_dims(self, axis):
        return ND(np.expand_dims(self.a, axis))

    def repeat(self, repeats, axis=None):
        return ND(np.repeat(self.a, repeats, axis=axis))

    def tile(self, reps):
        reps = (reps,) if isinstance(reps, int) else tuple(reps)
        return ND(np.tile(self.a, reps))

    def split(self, axis, num_outputs, squeeze_axis=False):
        parts = np.split(self.a, num_outputs, axis=axis)
        return [ND(np.squeeze(p, axis) if squeeze_axis else np.array(p)) for p in parts]

    def squeeze(self, axis=None):
        return ND(np.squeeze(self.a, axis=axis))

    def swapaxes(self, a1, a2):
        return ND(np.ascontiguousarray(np.swapaxes(self.a, a1, a2)))

    def argmax(self, axis):
        return ND(np.argmax(self.a, axis=axis).astype(f32))          # first maximum; MXNet returns fp32 indices

    def max(self, axis=None, keepdims=False):
        return ND(self.a.max(axis=axis, keepdims=keepdims))

    def clip(self, lo, hi):
        return ND(np.clip(self.a, f32(lo), f32(hi)))

    # ---- indexing (views write through, like NDArray slices)
    def __getitem__(self, k):
        r = self.a[k]
        return ND(r) if isinstance(r, np.ndarray) else ND(np.asarray(r, f32))

    def __setitem__(self, k, v):
        self.a[k] = np.asarray(_unwrap(v)).astype(f32)

    # ---- arithmetic (fp32)
    def _bin(self, o, fn, rev=False):
        o = _unwrap(o)
        o = np.asarray(o, f32) if not isinstance(o, np.ndarray) else o.astype(f32, copy=False)
        return ND(fn(o, self.a).astype(f32) if rev else fn(self.a, o).astype(f32))

    __add__ = lambda s, o: s._bin(o, np.add); __radd__ = __add__
    __mul__ = lambda s, o: s._bin(o, np.multiply); __rmul__ = __mul__
    __sub__ = lambda s, o: s._bin(o, np.subtract); __rsub__ = lambda s, o: s._bin(o, np.subtract, True)
    __truediv__ = lambda s, o: s._bin(o, np.divide); __rtruediv__ = lambda s, o: s._bin(o, np.divide, True)
    __neg__ = lambda s: ND(-s.a)
    __gt__ = lambda s, o: s._bin(o, lambda a, b: (a > b)); __ge__ = lambda s, o: s._bin(o, lambda a, b: (a >= b))
    __lt__ = lambda s, o: s._bin(o, lambda a, b: (a < b)); __le__ = lambda s, o: s._bin(o, lambda a, b: (a <= b))


class _Contrib:
    @staticmethod
    def box_nms(data, **kw):
        """mx.nd.contrib.box_nms: delegated to the oracle's restatement (oracle/ref_nms.py) -- what the goldens pin around it
        is the caller's wiring (concat order, parameters, slicing), not the operator."""
        import os, sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
        from oracle import ref_nms
        return ND(np.asarray(ref_nms.box_nms(data.a, **kw), f32))

    @staticmethod
    def box_iou(lhs, rhs, format="corner"):
        """mx.nd.contrib.box_iou, corner format: out[lhs..., rhs...] ; w,h clamped at 0 ; 0 where the union is <= 0."""
        a, b = lhs.a.reshape(-1, 4), rhs.a.reshape(-1, 4)
        l = np.maximum(a[:, None, 0], b[None, :, 0]); t = np.maximum(a[:, None, 1], b[None, :, 1])
        r = np.minimum(a[:, None, 2], b[None, :, 2]); bt = np.minimum(a[:, None, 3], b[None, :, 3])
        w = np.maximum(f32(0), r - l); h = np.maximum(f32(0), bt - t)
        i = (w * h).astype(f32)
        u = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None] + ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :] - i
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.where(u <= 0, f32(0), i / u).astype(f32)
        return ND(out.reshape(lhs.shape[:-1] + rhs.shape[:-1]))


class F:
    """The `F` / `nd` namespace."""
    contrib = _Contrib
    concat = staticmethod(lambda *xs, dim=1: ND(np.concatenate([x.a for x in xs], axis=dim)))
    zeros_like = staticmethod(lambda x: ND(np.zeros_like(x.a)))
    ones_like = staticmethod(lambda x: ND(np.ones_like(x.a)))
    sigmoid = staticmethod(lambda x: ND((f32(1) / (f32(1) + np.exp(-x.a))).astype(f32)))
    exp = staticmethod(lambda x: ND(np.exp(x.a).astype(f32)))
    relu = staticmethod(lambda x: ND(np.maximum(x.a, f32(0))))
    broadcast_add = staticmethod(lambda a, b: ND((a.a + b.a).astype(f32)))
    broadcast_mul = staticmethod(lambda a, b: ND((a.a * b.a).astype(f32)))
    broadcast_maximum = staticmethod(lambda a, b: ND(np.maximum(a.a, b.a)))
    broadcast_minimum = staticmethod(lambda a, b: ND(np.minimum(a.a, b.a)))
    arange = staticmethod(lambda start, stop=None: ND(np.arange(start, stop, dtype=f32)))
    tile = staticmethod(lambda x, reps: x.tile(reps))
    transpose = staticmethod(lambda x, axes=None: x.transpose(axes))
    reshape = staticmethod(lambda x, shape: x.reshape(shape))
    where = staticmethod(lambda c, a, b: ND(np.where(c.a != 0, a.a, b.a).astype(f32)))
    stop_gradient = staticmethod(lambda x: x)
    split = staticmethod(lambda x, axis, num_outputs, squeeze_axis=False: x.split(axis, num_outputs, squeeze_axis))

    max = staticmethod(lambda x, axis=None, keepdims=False: x.max(axis=axis, keepdims=keepdims))
    mean = staticmethod(lambda x, axis=None, keepdims=False: ND(x.a.mean(axis=axis, keepdims=keepdims, dtype=f32)))
    squeeze = staticmethod(lambda x, axis=None: x.squeeze(axis=axis))
    swapaxes = staticmethod(lambda x, dim1=0, dim2=0: x.swapaxes(dim1, dim2))
    repeat = staticmethod(lambda x, repeats=1, axis=None: x.repeat(repeats, axis=axis))
    expand_dims = staticmethod(lambda x, axis: x.expand_dims(axis))
    slice_axis = staticmethod(lambda x, axis, begin, end: x.slice_axis(axis, begin, end))

    @staticmethod
    def reshape_like(lhs, rhs, lhs_begin=None, lhs_end=None, rhs_begin=None, rhs_end=None):
        ls, rs = list(lhs.shape), list(rhs.shape)
        lb, le = (0 if lhs_begin is None else lhs_begin), (len(ls) if lhs_end is None else lhs_end)
        rb, re_ = (0 if rhs_begin is None else rhs_begin), (len(rs) if rhs_end is None else rhs_end)
        return ND(lhs.a.reshape(ls[:lb] + rs[rb:re_] + ls[le:]))

    @staticmethod
    def slice_like(x, like, axes):
        sl = [slice(None)] * x.a.ndim
        for ax in axes:
            sl[ax] = slice(0, like.a.shape[ax])
        return ND(np.array(x.a[tuple(sl)]))

    @staticmethod
    def one_hot(idx, depth):
        i = idx.a.astype(np.int64)
        out = np.zeros(i.shape + (depth,), f32)
        np.put_along_axis(out, i[..., None], f32(1), axis=-1)
        return ND(out)


class autograd:
    training = False
    pause = staticmethod(contextlib.nullcontext)
    is_training = staticmethod(lambda: autograd.training)
    is_recording = staticmethod(lambda: False)


class _Params:
    def get_constant(self, name, value):
        nd_ = ND(np.asarray(value, f32))
        nd_._is_param = True                     # Gluon hands registered parameters to hybrid_forward as keyword arguments
        return nd_


class _Block:
    def __init__(self, **kwargs):
        self.params = _Params()

    def name_scope(self):
        return contextlib.nullcontext()

    def __call__(self, *args):
        if hasattr(self, "hybrid_forward"):
            params = {k: v for k, v in self.__dict__.items() if isinstance(v, ND) and getattr(v, "_is_param", False)}
            return self.hybrid_forward(F, *args, **params)
        return self.forward(*args)


class gluon:
    Block = _Block
    HybridBlock = _Block


PARAM_LOG = []          # parameters of every shim layer in order of first execution (the generator exports them)
PARAM_RNG = np.random.RandomState(0)


def _bf16_round(x):
    u = np.ascontiguousarray(x, f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(f32).reshape(np.shape(x))


class _Conv2D:
    """nn.Conv2D: fp32 cross-correlation, stride 1, zero padding, optional bias.  Weights: set by the caller (`weight`, `bias`)
    or drawn at the first call (in_channels is inferred then, like Gluon's deferred init) and appended to PARAM_LOG."""

    def __init__(self, channels, kernel_size=1, strides=1, padding=0, use_bias=True, in_channels=0, **kw):
        k = kernel_size if isinstance(kernel_size, int) else kernel_size[0]
        p_ = padding if isinstance(padding, int) else padding[0]
        s_ = strides if isinstance(strides, int) else strides[0]
        assert s_ == 1 and p_ == k // 2, "shim conv: stride 1, 'same' padding only"
        self.channels, self.k, self.use_bias, self.weight, self.bias = channels, k, use_bias, None, None

    def __call__(self, x):
        cin = x.shape[1]
        if self.weight is None:
            fan = cin * self.k * self.k
            self.weight = _bf16_round((PARAM_RNG.standard_normal((self.channels, cin, self.k, self.k)) * np.sqrt(2.0 / fan)).astype(f32))
            self.bias = PARAM_RNG.uniform(-0.2, 0.2, self.channels).astype(f32) if self.use_bias else None
            PARAM_LOG.append(("conv", self.weight, self.bias))
        w = np.asarray(self.weight, f32).reshape(self.channels, cin, self.k, self.k)
        B, _, H, W = x.shape
        p_ = self.k // 2
        xp = np.zeros((B, cin, H + 2 * p_, W + 2 * p_), f32)
        xp[:, :, p_:p_ + H, p_:p_ + W] = x.a
        y = np.zeros((B, self.channels, H, W), f32)
        for j in range(self.k):
            for i in range(self.k):
                y += np.einsum("nk,bkhw->bnhw", w[:, :, j, i], xp[:, :, j:j + H, i:i + W], optimize=True).astype(f32)
        if self.use_bias and self.bias is not None:
            y = y + np.asarray(self.bias, f32).reshape(1, -1, 1, 1)
        return ND(y.astype(f32))


class _Conv3D:
    """nn.Conv3D on (B, C, T, H, W): fp32 cross-correlation, stride 1, zero 'same' padding, no bias; kernel extents 1 or 3.
    Weights (Co, Ci, kt, kh, kw) drawn at the first call and logged like _Conv2D."""

    def __init__(self, channels, kernel_size=1, strides=1, padding=0, use_bias=True, groups=1, **kw):
        k = (kernel_size,) * 3 if isinstance(kernel_size, int) else tuple(kernel_size)
        p_ = (padding,) * 3 if isinstance(padding, int) else tuple(padding)
        s_ = (strides,) * 3 if isinstance(strides, int) else tuple(strides)
        assert s_ == (1, 1, 1) and p_ == tuple(e // 2 for e in k) and not use_bias and groups == 1, "shim conv3d: stride 1, 'same', no bias"
        self.channels, self.k, self.weight = channels, k, None

    def __call__(self, x):
        cin = x.shape[1]
        kt, kh, kw = self.k
        if self.weight is None:
            fan = cin * kt * kh * kw
            self.weight = _bf16_round((PARAM_RNG.standard_normal((self.channels, cin, kt, kh, kw)) * np.sqrt(2.0 / fan)).astype(f32))
            PARAM_LOG.append(("conv", self.weight, None))
        B, _, T, H, W = x.shape
        pt, ph, pw = kt // 2, kh // 2, kw // 2
        xp = np.zeros((B, cin, T + 2 * pt, H + 2 * ph, W + 2 * pw), f32)
        xp[:, :, pt:pt + T, ph:ph + H, pw:pw + W] = x.a
        y = np.zeros((B, self.channels, T, H, W), f32)
        for i in range(kt):
            for j in range(kh):
                for k in range(kw):
                    y += np.einsum("nk,bkthw->bnthw", self.weight[:, :, i, j, k], xp[:, :, i:i + T, j:j + H, k:k + W], optimize=True).astype(f32)
        return ND(y.astype(f32))


class BatchNorm:
    """Inference BatchNorm over axis 1: (x - mean) / sqrt(var + eps) * gamma + beta (fp32); statistics drawn at the first call."""

    def __init__(self, epsilon=1e-5, momentum=0.9, **kw):
        self.eps, self.p = epsilon, None

    def __call__(self, x):
        c = x.shape[1]
        if self.p is None:
            self.p = (PARAM_RNG.uniform(0.5, 1.5, c).astype(f32), PARAM_RNG.uniform(-0.2, 0.2, c).astype(f32),
                      PARAM_RNG.uniform(-0.2, 0.2, c).astype(f32), PARAM_RNG.uniform(0.5, 1.5, c).astype(f32))
            PARAM_LOG.append(("bn",) + self.p)
        g, b, m, v = [t.reshape((1, c) + (1,) * (x.ndim - 2)) for t in self.p]
        return ND(((x.a - m) / np.sqrt(v + f32(self.eps)) * g + b).astype(f32))


class _LeakyReLU:
    def __init__(self, alpha):
        self.alpha = f32(alpha)

    def __call__(self, x):
        return ND(np.where(x.a > 0, x.a, x.a * self.alpha).astype(f32))


class _HybridSequential:
    def __init__(self, prefix=None, **kw):
        self.layers = []

    def add(self, *blocks):
        self.layers.extend(blocks)

    def __call__(self, x):
        for l in self.layers:
            x = l(x)
        return x

    def __getitem__(self, i):
        return self.layers[i]

    def __len__(self):
        return len(self.layers)

    def __iter__(self):
        return iter(self.layers)


class nn:
    Conv2D = _Conv2D
    Conv3D = _Conv3D
    BatchNorm = BatchNorm
    LeakyReLU = _LeakyReLU
    HybridSequential = _HybridSequential


# ---- gluoncv.nn.bbox (published definitions, gluon-cv 0.4/0.5)
class BBoxCornerToCenter:
    def __init__(self, axis=-1, split=False):
        self.axis, self.split_out = axis, split

    def __call__(self, x):
        xmin, ymin, xmax, ymax = x.split(axis=self.axis, num_outputs=4)
        width, height = xmax - xmin, ymax - ymin
        cx, cy = xmin + width / 2, ymin + height / 2
        return (cx, cy, width, height) if self.split_out else F.concat(cx, cy, width, height, dim=self.axis)


class BBoxCenterToCorner:
    def __init__(self, axis=-1, split=False):
        self.axis, self.split_out = axis, split

    def __call__(self, x):
        cx, cy, w, h = x.split(axis=self.axis, num_outputs=4)
        hw, hh = w / 2, h / 2
        out = (cx - hw, cy - hh, cx + hw, cy + hh)
        return out if self.split_out else F.concat(*out, dim=self.axis)


class BBoxBatchIOU:
    """gluoncv.nn.bbox.BBoxBatchIOU(axis=-1, fmt='corner', offset=0, eps=1e-15): (B,N,4) x (B,M,4) -> (B,N,M)."""

    def __init__(self, axis=-1, fmt="corner", offset=0, eps=1e-15):
        self.offset, self.eps = offset, eps

    def __call__(self, a, b):
        al, at, ar, ab = a.split(axis=-1, num_outputs=4, squeeze_axis=True)      # (B,N)
        bl, bt, br, bb = b.split(axis=-1, num_outputs=4, squeeze_axis=True)      # (B,M)
        left = F.broadcast_maximum(al.expand_dims(-1), bl.expand_dims(-2))
        right = F.broadcast_minimum(ar.expand_dims(-1), br.expand_dims(-2))
        top = F.broadcast_maximum(at.expand_dims(-1), bt.expand_dims(-2))
        bot = F.broadcast_minimum(ab.expand_dims(-1), bb.expand_dims(-2))
        iw = F.relu(right - left + self.offset)
        ih = F.relu(bot - top + self.offset)
        i = iw * ih
        area_a = ((ar - al + self.offset) * (ab - at + self.offset)).expand_dims(-1)
        area_b = ((br - bl + self.offset) * (bb - bt + self.offset)).expand_dims(-2)
        union = F.broadcast_add(area_a, area_b) - i
        return i / (union + self.eps)


def namespace():
    """Globals for exec'ing the reference classes."""
    import warnings
    import math
    return {"np": np, "math": math, "nd": F, "gluon": gluon, "nn": nn, "autograd": autograd, "warnings": warnings, "BatchNorm": BatchNorm,
            "YOLOV3Loss": type("YOLOV3Loss", (), {}),
            "BBoxCornerToCenter": BBoxCornerToCenter, "BBoxCenterToCorner": BBoxCenterToCorner, "BBoxBatchIOU": BBoxBatchIOU}
