"""On-disk feature format -> head carriers.

`extract_base_features.py:153-155` writes, per frame, three fp32 NCHW arrays `<file_id>_F1.npy` (256 ch, stride 8),
`_F2.npy` (512 ch, stride 16), `_F3.npy` (1024 ch, stride 32); the datasets read them back with `np.load`
(`datasets/pascalvoc.py:112-114`, `datasets/imgnetvid.py:155-157,182-184`) and `YOLOV3_noback` consumes them
(`yolo3.py:1784`).  `FeatureStream` streams such files into the channels-last bf16 carriers the fused head takes:
memory-mapped reads -> pinned fp32 staging -> async H2D on a copy stream -> vd_repack_nchw_f32_to_nhwc_bf16,
double-buffered so batch i+1 is loaded while batch i is being processed.  Output order is the head's: [s32, s16, s8]
= [F3, F2, F1].  Host code only orchestrates; there is no CPU compute path for the head itself.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ._lib import check, load, ptr

SUFFIXES_OUT_ORDER = ("_F3.npy", "_F2.npy", "_F1.npy")      # s32, s16, s8


def feature_paths(features_dir, file_id):
    """The three files of one frame in the head's scale order (s32, s16, s8)."""
    return [os.path.join(features_dir, file_id + s) for s in SUFFIXES_OUT_ORDER]


def batches(file_ids, batch):
    """Contiguous batches of file ids (the last one may be short)."""
    return [list(file_ids[i:i + batch]) for i in range(0, len(file_ids), batch)]


class FeatureStream:
    """Iterates over batches of frames stored as `<file_id>_F{1,2,3}.npy`; yields `(ids, [s32, s16, s8])` with each tip a
    channels-last bf16 CUDA tensor (n, C, H, W).  `window=T` groups T consecutive ids per sample: tips are (n, T, C, H, W)."""

    def __init__(self, features_dir, file_ids, batch, device="cuda", window=None, depth=2):
        self.features_dir, self.file_ids, self.batch = features_dir, list(file_ids), int(batch)
        self.device = torch.device(device)
        self.window = window
        self.depth = max(2, int(depth))
        if window is not None:
            assert len(self.file_ids) % window == 0, "file ids must come in whole windows"
        first = [np.load(p, mmap_mode="r") for p in feature_paths(features_dir, self.file_ids[0])]
        for a in first:
            if a.dtype != np.float32 or a.ndim != 3:
                raise ValueError("feature files must hold fp32 (C,H,W) arrays, got %s %s" % (a.dtype, a.shape))
        self.shapes = [tuple(a.shape) for a in first]
        per = self.batch * (window or 1)
        self._frames_per_batch = per
        self._copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._slots = None

    def _alloc(self):
        per = self._frames_per_batch
        slots = []
        for _ in range(self.depth):
            host = [torch.empty((per,) + shp, dtype=torch.float32).pin_memory() for shp in self.shapes]
            dev32 = [torch.empty((per,) + shp, dtype=torch.float32, device=self.device) for shp in self.shapes]
            out = [torch.empty((per,) + shp, dtype=torch.bfloat16, device=self.device).contiguous(memory_format=torch.channels_last)
                   for shp in self.shapes]
            slots.append({"host": host, "dev32": dev32, "out": out, "ready": torch.cuda.Event(), "free": None})
        self._slots = slots

    def _stage(self, slot, ids):
        """Disk -> pinned staging (host work), then async H2D + repack on the copy stream."""
        if slot["free"] is not None:
            slot["free"].synchronize()                      # the consumer of this slot's previous batch has finished
        n = len(ids)
        for k in range(3):
            h = slot["host"][k].numpy()
            for i, fid in enumerate(ids):
                a = np.load(os.path.join(self.features_dir, fid + SUFFIXES_OUT_ORDER[k]), mmap_mode="r")
                if tuple(a.shape) != self.shapes[k]:
                    raise ValueError("%s%s has shape %s, expected %s" % (fid, SUFFIXES_OUT_ORDER[k], a.shape, self.shapes[k]))
                h[i] = a
        lib = load()
        with torch.cuda.stream(self._copy_stream):
            for k in range(3):
                C, H, W = self.shapes[k]
                slot["dev32"][k][:n].copy_(slot["host"][k][:n], non_blocking=True)
                check(lib.vd_repack_nchw_f32_to_nhwc_bf16(ptr(slot["dev32"][k]), ptr(slot["out"][k]), n, C, H, W,
                                                         torch.cuda.current_stream().cuda_stream))
            slot["ready"].record(self._copy_stream)
        return n

    def __iter__(self):
        if self.device.type != "cuda":
            raise RuntimeError("FeatureStream feeds the CUDA head: there is no CPU path")
        if self._slots is None:
            self._alloc()
        groups = batches(self.file_ids, self._frames_per_batch)
        pending = []
        nxt = 0
        for slot in self._slots[: min(self.depth, len(groups))]:
            pending.append((slot, groups[nxt], self._stage(slot, groups[nxt]))); nxt += 1
        while pending:
            slot, ids, n = pending.pop(0)
            torch.cuda.current_stream().wait_event(slot["ready"])
            tips = [o[:n] for o in slot["out"]]
            if self.window is not None:
                tips = [t.reshape((n // self.window, self.window) + tuple(t.shape[1:])) for t in tips]
            yield ids, tips
            slot["free"] = torch.cuda.Event()
            slot["free"].record(torch.cuda.current_stream())
            if nxt < len(groups):
                pending.append((slot, groups[nxt], self._stage(slot, groups[nxt]))); nxt += 1
