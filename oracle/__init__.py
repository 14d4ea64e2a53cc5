"""CPU oracle for the VidDet head -> decode -> NMS (+ target generation) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``viddet_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and there only as the checker / the timed CPU baseline.

PARITY STATUS: MXNet / GluonCV (``mxnet-cu100``, ``gluoncv``; requirements.txt:1-2, un-vendored and un-pinned) cannot be
imported or built here and HaydenFaulkner/VidDet ships no tests, golden vectors or fixtures, so the reference cannot be run
end to end.  What IS pinned, by executing the reference's own source (cut out of /root/reference with ``ast`` at generation
time, never copied):
  * ``hierarchical_nms`` + ``iou`` + the CombinedDetection tree methods (pure Python): tests/golden/hier_nms_golden.npz
    (tests/golden/make_golden_hier_nms.py) -- fully pinned, bit-exact;
  * ``YOLOOutputV3.hybrid_forward`` (all three modes), ``YOLOV3PrefetchTargetGenerator.forward/_slice``,
    ``YOLOV3DynamicTargetGeneratorSimple`` and ``YOLOV3TargetMerger``: tests/golden/ref_exec_golden.npz
    (tests/golden/make_golden_ref_exec.py), executed over tests/golden/mx_shim.py, a numpy stand-in for the MXNet / GluonCV operators
    they call -- pins the reference's own logic (slicing, reshape/transpose row order, the per-GT loop, index math, _slice,
    where-merges); the operators inside the shim are restated from their published definitions;
  * ``utils/bbox.py::bbox_iou`` (importable numpy): tests/golden/bbox_iou_golden.npz (tests/golden/make_golden_bbox_iou.py);
  * the worked examples of the upstream ``box_nms`` / ``box_iou`` operator documentation (tests/test_oracle_kat.py).
Still **parity unpinned** (restated from published algorithms only): the MXNet operators themselves
-- ``contrib.box_nms`` (sort / top-k / greedy suppression), ``contrib.box_iou``, ``Convolution``, ``BatchNorm``, and GluonCV's
``YOLOV3Loss``.
"""
