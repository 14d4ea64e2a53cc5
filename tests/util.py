"""Shared synthetic-input builders for the parity tests (SURVEY.md section 8d distributions)."""
import numpy as np

ANCHORS = [[116, 90, 156, 198, 373, 326], [30, 61, 62, 45, 59, 119], [10, 13, 16, 30, 33, 23]]
STRIDES = [32, 16, 8]
CHANNELS = [1024, 512, 256]


def bf16_round(x):
    """fp32 -> nearest-even bf16 -> fp32 (numpy)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def make_tips(rng, B, size=416, channels=CHANNELS, strides=STRIDES, T=None):
    """tips = leaky_relu(N(0,1), 0.1), NCHW fp32, already rounded to bf16-representable values."""
    tips = []
    for c, s in zip(channels, strides):
        h = size // s
        shape = (B, c, h, h) if T is None else (B, T, c, h, h)
        x = rng.standard_normal(shape).astype(np.float32)
        x = np.where(x > 0, x, 0.1 * x).astype(np.float32)
        tips.append(bf16_round(x))
    return tips


def make_pred_weights(rng, num_class, channels=CHANNELS, scale=0.07, bias_scale=0.0, k=1):
    ws, bs = [], []
    n = 3 * (5 + num_class)
    for c in channels:
        ws.append(bf16_round(rng.uniform(-scale, scale, size=(n, c * k, 1, 1)).astype(np.float32)))
        bs.append((rng.uniform(-bias_scale, bias_scale, size=(n,)) if bias_scale else np.zeros(n)).astype(np.float32))
    return ws, bs


def random_dets(rng, B, N, num_class=5, size=416.0, tie_frac=0.0, invalid_frac=0.1):
    """(B,N,6) rows [id, score, x1,y1,x2,y2] with clustered boxes so NMS has work to do."""
    ids = rng.randint(0, num_class, size=(B, N)).astype(np.float32)
    scores = rng.uniform(0.0, 1.0, size=(B, N)).astype(np.float32)
    if tie_frac > 0:
        m = rng.uniform(size=(B, N)) < tie_frac
        scores[m] = np.round(scores[m] * 8) / 8          # many exact ties
    inv = rng.uniform(size=(B, N)) < invalid_frac
    scores[inv] = rng.uniform(-0.5, 0.01, size=inv.sum()).astype(np.float32)
    ncl = max(4, N // 6)
    centers = rng.uniform(0.1 * size, 0.9 * size, size=(B, ncl, 2))
    which = rng.randint(0, ncl, size=(B, N))
    c = np.take_along_axis(centers, which[..., None].repeat(2, -1), axis=1) + rng.normal(0, 6.0, size=(B, N, 2))
    wh = np.exp(rng.uniform(np.log(20), np.log(120), size=(B, N, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], -1)
    return np.concatenate([ids[..., None], scores[..., None], boxes], -1).astype(np.float32)


def make_gt(rng, B, M, size=416, num_class=20, multi_hot=False, min_count=0):
    """gt_boxes (B,M,4) corner px padded with -1, gt_ids (B,M,1) or multi-hot (B,M,C)."""
    gt = np.full((B, M, 4), -1.0, np.float32)
    ids = np.full((B, M, num_class if multi_hot else 1), -1.0 if not multi_hot else 0.0, np.float32)
    for b in range(B):
        n = rng.randint(min_count, M + 1)
        w = np.exp(rng.uniform(np.log(8), np.log(min(400, size - 2)), size=n))
        h = np.exp(rng.uniform(np.log(8), np.log(min(400, size - 2)), size=n))
        cx = rng.uniform(w / 2, size - w / 2)
        cy = rng.uniform(h / 2, size - h / 2)
        # keep centres off exact grid lines (A.4: fp64-vs-fp32 only differs within 1 ulp of an integer)
        cx = np.floor(cx) + 0.37
        cy = np.floor(cy) + 0.61
        x1, y1 = np.maximum(cx - w / 2, 0), np.maximum(cy - h / 2, 0)
        gt[b, :n] = np.stack([x1, y1, np.minimum(cx + w / 2, size - 1e-3), np.minimum(cy + h / 2, size - 1e-3)], -1)
        if multi_hot:
            for m in range(n):
                k = rng.randint(1, 6)
                ids[b, m, rng.choice(num_class, size=k, replace=False)] = 1.0
        else:
            ids[b, :n, 0] = rng.randint(0, num_class, size=n)
    return gt, ids


def replay_shim_params(seed, kinds, shapes, has_bias):
    """Re-draws the layer parameters tests/golden/mx_shim.py drew (same RandomState stream, same expressions, same order) when the
    reference's YOLOV3.hybrid_forward was executed for tests/golden/ref_exec_golden.npz.  Returns a list of
    ('conv', weight, bias_or_None) / ('bn', gamma, beta, mean, var)."""
    rng = np.random.RandomState(seed)
    out = []
    for kind, shp, hb in zip(kinds, shapes, has_bias):
        if int(kind) == 0:
            shp = tuple(int(v) for v in shp if int(v) > 0)        # (Co, Ci, kh, kw) or (Co, Ci, kt, kh, kw); zero-padded in the file
            fan = int(np.prod(shp[1:]))
            w = bf16_round((rng.standard_normal(shp) * np.sqrt(2.0 / fan)).astype(np.float32))
            b = rng.uniform(-0.2, 0.2, shp[0]).astype(np.float32) if int(hb) else None
            out.append(("conv", w, b))
        else:
            c = int(shp[0])
            out.append(("bn", rng.uniform(0.5, 1.5, c).astype(np.float32), rng.uniform(-0.2, 0.2, c).astype(np.float32),
                        rng.uniform(-0.2, 0.2, c).astype(np.float32), rng.uniform(0.5, 1.5, c).astype(np.float32)))
    return out


def keep_agreement(keep_dev, rec_ref, det_ref, tol):
    """Compare the device's NMS keep rows (frames, post) with the oracle's record (frames, >= post) position by position.
    A mismatch is accepted only as a near-tie: the row the device kept at that position must have an ORACLE score within `tol`
    (relative) of the oracle's own pick there (two candidates whose scores differ by less than the arithmetic noise may swap
    ranks).  Returns (fraction of identical positions, number of near-tie positions); raises on any other mismatch."""
    post = keep_dev.shape[1]
    same = keep_dev == rec_ref[:, :post]
    n_tie = 0
    for f, i in zip(*np.nonzero(~same)):
        rd, rr = int(keep_dev[f, i]), int(rec_ref[f, i])
        assert rd >= 0 and rr >= 0, "frame %d pos %d: device kept row %d, oracle row %d" % (f, i, rd, rr)
        sd, sr = float(det_ref[f, rd, 1]), float(det_ref[f, rr, 1])
        assert abs(sd - sr) <= tol * abs(sr), "frame %d pos %d: rows %d / %d differ by more than a near-tie (%.9g vs %.9g)" % (f, i, rd, rr, sd, sr)
        n_tie += 1
    return float(same.mean()), n_tie


def rel_err(got, ref, floor):
    """max |got - ref| / max(|ref|, floor) over all elements."""
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float((np.abs(got - ref) / np.maximum(np.abs(ref), floor)).max())


def err_profile(got, ref, what):
    """(max, 99.9th percentile, mean) of |got - ref| relative to the largest |ref|; printed so the bounds below stay measured."""
    e = np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64)).ravel() / float(np.abs(ref).max())
    mx, p999, mean = float(e.max()), float(np.quantile(e, 0.999)), float(e.mean())
    print("%s: max %.2e, p99.9 %.2e, mean %.2e of max|ref|" % (what, mx, p999, mean))
    return mx, p999, mean
