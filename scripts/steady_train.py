"""A few calls of the training-side kernels at the cfg-5 shapes (B=128, C=285, M<=100): the target of `ncu -k regex:...`.
usage: python scripts/steady_train.py [calls]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import viddet_b200
from tests.util import ANCHORS, make_gt

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
B, M, C, size = 128, 100, 285, 416
rng = np.random.RandomState(1234)
gt, ids = make_gt(rng, B, M, size=size, num_class=C, multi_hot=True)
gb, gi = torch.from_numpy(gt).to(dev), torch.from_numpy(ids).to(dev)
gen = viddet_b200.YOLOV3PrefetchTargetGenerator(C)
hs = [size // s for s in bench.STRIDES]
xs = [(B, 1, h, h) for h in hs]
anchors = [np.asarray(a, np.float32).reshape(1, 1, 3, 2) for a in ANCHORS]
offsets = [np.zeros((1, h * h, 1, 2), np.float32) for h in hs]
n_anch = 3 * sum(h * h for h in hs)
outs = gen.alloc_outputs(B, n_anch, dev)
g = torch.Generator(device=dev).manual_seed(7)
box = torch.rand((B, n_anch, 4), generator=g, device=dev) * 300
box[..., 2:] += box[..., :2] + 8
mg = viddet_b200.YOLOV3TargetMerger(C, 0.7)
loss = viddet_b200.YOLOV3Loss()
preds = [torch.randn((B, n_anch, w), generator=g, device=dev) for w in (1, 2, 2, C)]
for _ in range(n):
    pre = gen.run_into((B, 3, size, size), xs, anchors, offsets, gb, gi, None, outs)
    merged = mg(box, gb, *pre)
    loss(*preds, *merged)
torch.cuda.synchronize()
print("done")
