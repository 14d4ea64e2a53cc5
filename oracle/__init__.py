"""CPU oracle for the VidDet head -> decode -> NMS (+ target generation) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``viddet_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU baseline.

PARITY STATUS: **parity unpinned by the reference** -- HaydenFaulkner/VidDet ships no tests,
golden vectors or fixtures, and its arithmetic lives in un-vendored, un-pinned third-party
packages (``mxnet-cu100``, ``gluoncv``; requirements.txt:1-2) that cannot be imported or built
here.  The oracle restates (a) the reference's own Python (file:line cited per function) and
(b) the published algorithms of the MXNet operators it calls.  What
*is* pinned: the two worked examples from the upstream ``box_nms`` / ``box_iou`` operator
documentation (tests/test_oracle_kat.py) and IoU values produced by the reference's own
importable numpy helper ``utils/bbox.py::bbox_iou`` (tests/golden/bbox_iou_golden.npz, made by
scripts/make_golden_bbox_iou.py).
"""
