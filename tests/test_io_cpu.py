"""Host-side logic of the feature-file loader (no GPU): naming, scale order, batching, input validation."""
import numpy as np
import pytest

from viddet_b200 import io


def test_paths_follow_the_reference_naming_and_head_scale_order(tmp_path):
    p = io.feature_paths(str(tmp_path), "ILSVRC2015_val_00000000/000010")
    assert [x.split("_")[-1] for x in p] == ["F3.npy", "F2.npy", "F1.npy"]       # s32, s16, s8
    assert io.batches(list("abcdefg"), 3) == [["a", "b", "c"], ["d", "e", "f"], ["g"]]


def test_rejects_wrong_dtype_and_needs_cuda(tmp_path):
    for suf, shp in zip(("_F1.npy", "_F2.npy", "_F3.npy"), ((4, 2, 2), (8, 1, 1), (16, 1, 1))):
        np.save(str(tmp_path / ("x" + suf)), np.zeros(shp, np.float64))
    with pytest.raises(ValueError):
        io.FeatureStream(str(tmp_path), ["x"], batch=1, device="cpu")
    for suf, shp in zip(("_F1.npy", "_F2.npy", "_F3.npy"), ((4, 2, 2), (8, 1, 1), (16, 1, 1))):
        np.save(str(tmp_path / ("y" + suf)), np.zeros(shp, np.float32))
    fs = io.FeatureStream(str(tmp_path), ["y"], batch=1, device="cpu")
    assert fs.shapes == [(16, 1, 1), (8, 1, 1), (4, 2, 2)]
    with pytest.raises(RuntimeError):
        next(iter(fs))                                   # no CPU path for the head
    with pytest.raises(AssertionError):
        io.FeatureStream(str(tmp_path), ["y"], batch=1, device="cpu", window=2)
