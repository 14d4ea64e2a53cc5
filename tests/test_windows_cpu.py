"""viddet_b200.window_frame_indices against golden vectors produced by executing the reference's own window construction
(datasets/imgnetvid.py:480-506; tests/golden/make_golden_windows.py)."""
import os

import numpy as np

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "windows_golden.npz"))


def test_window_frame_indices_match_the_executed_reference():
    from viddet_b200.blocks import window_frame_indices
    flat, pos = G["flat"], 0
    for L, T, step, wlen in G["cases"]:
        for centre in range(L):
            exp = flat[pos:pos + wlen].tolist()
            pos += wlen
            assert window_frame_indices(int(L), centre, int(T), int(step)) == exp, (L, T, step, centre)
    assert pos == len(flat)


def test_sliding_windows_are_the_unclamped_case():
    """ClipWindows covers consecutive centres away from the clip's ends at step 1: window b = frames [b, b+T)."""
    from viddet_b200.blocks import window_frame_indices
    L, T = 40, 5
    for b in range(L - T + 1):
        assert window_frame_indices(L, b + T // 2, T) == list(range(b, b + T))
