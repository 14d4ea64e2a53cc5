"""CPU checks of the conv-BN-LReLU / YOLODetectionBlockV3 oracle (oracle/ref_block.py): hand-derived KATs and an
independent restatement (torch conv3d, fp32) -- the reference itself (MXNet) cannot run, parity is unpinned."""
import numpy as np
import torch

from oracle import ref_block, ref_temporal

f32 = np.float32


def lrelu(a):
    return np.where(a > 0, a, 0.1 * a).astype(f32)


def ident(c):
    return np.ones(c, f32), np.zeros(c, f32), np.zeros(c, f32), np.ones(c, f32) - f32(1e-5)


def test_shift_taps_and_zero_padding():
    rng = np.random.RandomState(0)
    B, T, C, H, W = 1, 3, 4, 3, 5
    x = rng.standard_normal((B, T, C, H, W)).astype(f32)
    # tap (it=1, iy=1, ix=2): y[.., x] = x[.., x+1], zero in the last column (cross-correlation, layers.py:66)
    w = np.zeros((C, C, 3, 3, 3), f32); w[:, :, 1, 1, 2] = np.eye(C)
    y = ref_block.conv_bn_lrelu(x, w, *ident(C))
    exp = np.zeros_like(x); exp[..., :-1] = x[..., 1:]
    np.testing.assert_allclose(y, lrelu(exp), rtol=1e-6, atol=1e-7)
    # tap (it=0, iy=2, ix=1): y[t, y] = x[t-1, y+1]
    w = np.zeros((C, C, 3, 3, 3), f32); w[:, :, 0, 2, 1] = np.eye(C)
    y = ref_block.conv_bn_lrelu(x, w, *ident(C))
    exp = np.zeros_like(x); exp[:, 1:, :, :-1] = x[:, :-1, :, 1:]
    np.testing.assert_allclose(y, lrelu(exp), rtol=1e-6, atol=1e-7)


def test_pointwise_and_bn_fold():
    rng = np.random.RandomState(1)
    x = rng.standard_normal((2, 6, 4, 4)).astype(f32)
    w = rng.standard_normal((3, 6, 1, 1)).astype(f32)
    g, b, m, v = (rng.uniform(0.5, 1.5, 3).astype(f32), rng.uniform(-1, 1, 3).astype(f32),
                  rng.uniform(-1, 1, 3).astype(f32), rng.uniform(0.5, 1.5, 3).astype(f32))
    y = ref_block.conv_bn_lrelu(x, w, g, b, m, v)
    z = np.einsum("oc,bchw->bohw", w[:, :, 0, 0], x)
    z = (z - m.reshape(1, 3, 1, 1)) / np.sqrt(v.reshape(1, 3, 1, 1) + 1e-5) * g.reshape(1, 3, 1, 1) + b.reshape(1, 3, 1, 1)
    np.testing.assert_allclose(y, lrelu(z), rtol=1e-5, atol=1e-6)


def test_against_torch_conv3d_and_temporal_oracle():
    rng = np.random.RandomState(2)
    for k in [(1, 3, 3), (3, 1, 1), (3, 3, 3), (1, 1, 1)]:
        B, T, C, Co, H, W = 2, 4, 5, 7, 6, 5
        x = rng.standard_normal((B, T, C, H, W)).astype(f32)
        w = rng.standard_normal((Co, C) + k).astype(f32)
        y = ref_block.conv_bn_lrelu(x, w, *ident(Co))
        z = torch.nn.functional.conv3d(torch.from_numpy(x).permute(0, 2, 1, 3, 4), torch.from_numpy(w),
                                       padding=tuple(e // 2 for e in k)).permute(0, 2, 1, 3, 4).numpy()
        np.testing.assert_allclose(y, lrelu(z), rtol=1e-4, atol=1e-4)
    C = 6
    x = rng.standard_normal((1, 5, C, 3, 3)).astype(f32)
    w = rng.standard_normal((C, C, 3, 1, 1)).astype(f32)
    np.testing.assert_allclose(ref_block.conv_bn_lrelu(x, w, *ident(C)),
                               ref_temporal.temporal_conv_bn_lrelu(x, w, *ident(C)), rtol=1e-5, atol=1e-5)


def test_detection_block_wiring():
    """Body = 5 cells, tip = 1 cell (conv_type '2'); '21' doubles every expand (yolo3_temporal.py:204-227)."""
    rng = np.random.RandomState(3)
    ch, cin = 4, 6

    def cell(ci, co, k):
        return dict(weight=rng.uniform(-0.3, 0.3, (co, ci) + k).astype(f32), gamma=np.ones(co, f32), beta=np.zeros(co, f32),
                    mean=np.zeros(co, f32), var=np.ones(co, f32))
    cells = [cell(cin, ch, (1, 1)), cell(ch, 2 * ch, (3, 3)), cell(2 * ch, ch, (1, 1)), cell(ch, 2 * ch, (3, 3)),
             cell(2 * ch, ch, (1, 1)), cell(ch, 2 * ch, (3, 3))]
    x = rng.standard_normal((2, cin, 5, 5)).astype(f32)
    route, tip = ref_block.detection_block(x, cells, "2")
    assert route.shape == (2, ch, 5, 5) and tip.shape == (2, 2 * ch, 5, 5)
    z = x
    for c in cells[:5]:
        z = ref_block.conv_bn_lrelu(z, c["weight"], c["gamma"], c["beta"], c["mean"], c["var"])
    np.testing.assert_array_equal(route, z)
    cells21 = [cell(cin, ch, (1, 1, 1)), cell(ch, 2 * ch, (1, 3, 3)), cell(2 * ch, 2 * ch, (3, 1, 1)), cell(2 * ch, ch, (1, 1, 1)),
               cell(ch, 2 * ch, (1, 3, 3)), cell(2 * ch, 2 * ch, (3, 1, 1)), cell(2 * ch, ch, (1, 1, 1)),
               cell(ch, 2 * ch, (1, 3, 3)), cell(2 * ch, 2 * ch, (3, 1, 1))]
    x5 = rng.standard_normal((1, 3, cin, 4, 4)).astype(f32)
    route, tip = ref_block.detection_block(x5, cells21, "21")
    assert route.shape == (1, 3, ch, 4, 4) and tip.shape == (1, 3, 2 * ch, 4, 4)
