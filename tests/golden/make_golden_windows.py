"""Generates tests/golden/windows_golden.npz by EXECUTING the reference's own window construction
(datasets/imgnetvid.py:480-506, the `if self._window_size > 1:` block of the frame-sample loader): the lines are cut out of
/root/reference at generation time (never copied into the repo), dedented and run with a stand-in `self` and `videos`.
Run in the authoring container only; the .npz travels to the GPU box."""
import os
import textwrap
import types

import numpy as np

REF = "/root/reference/datasets/imgnetvid.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "windows_golden.npz")


def reference_windows(num_frames, window_size, window_step):
    lines = open(REF).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.strip() == "if self._window_size > 1:")
    end = next(i for i in range(start, len(lines)) if lines[i].strip().startswith("return frames"))
    block = textwrap.dedent("\n".join(lines[start:end]))
    me = types.SimpleNamespace(_window_size=window_size, _window_step=window_step, _windows=None)
    videos = {"v": ("val", "clip", ["%06d" % i for i in range(num_frames)], list(range(num_frames)))}
    exec(compile(block, REF, "exec"), {"self": me, "videos": videos})
    return [me._windows[i] for i in range(num_frames)] if me._windows is not None else [[i] for i in range(num_frames)]


def main():
    cases, flat = [], []
    for L in (1, 2, 3, 5, 9, 14):
        for T in (2, 3, 4, 5, 7):
            for step in (1, 2, 3):
                w = reference_windows(L, T, step)
                cases.append((L, T, step, len(w[0])))
                flat.extend(v for win in w for v in win)
    np.savez_compressed(OUT, cases=np.array(cases, np.int32), flat=np.array(flat, np.int32))
    print("wrote", OUT, len(cases), "cases,", len(flat), "indices")


if __name__ == "__main__":
    main()
