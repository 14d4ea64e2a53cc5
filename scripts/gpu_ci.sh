#!/bin/bash
# Runs the GPU test files one process each (a CUDA fault in one cannot poison the others), with a
# hard timeout per file; logs under gpurun_out/.  Usage (under gpurun): bash scripts/gpu_ci.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
files="$@"
if [ -z "$files" ]; then files="tests/test_gpu_nms.py tests/test_gpu_decode.py tests/test_gpu_targets.py tests/test_gpu_train.py tests/test_gpu_block.py tests/test_gpu_ref_exec.py tests/test_gpu_edges.py tests/test_gpu_head.py tests/test_gpu_fp32.py tests/test_gpu_peer.py"; fi
rc_all=0
for f in $files; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -q -s -m gpu > gpurun_out/$name.log 2>&1
  rc=$?
  echo "$f exit $rc" | tee -a gpurun_out/summary.txt
  grep -E 'err|identical|redone|passed|failed|FAILED|Error|error' gpurun_out/$name.log | tail -n 40
  if [ $rc -ne 0 ]; then rc_all=1; fi
done
exit $rc_all
