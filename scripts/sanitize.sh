#!/bin/bash
# compute-sanitizer over the head / NMS / pipeline / CTA-pair conv tests (summaries -> gpurun_out/sanitizer_*.txt; copy into profiles/).
# Usage (under gpurun): bash scripts/sanitize.sh
mkdir -p gpurun_out
SEL='test_fused_equals_compat_chain_bit_exact and (4-160-5 or 20-416-3) or test_pipeline_matches_serial_calls or test_output_mirrors or test_arbitrary_class_counts_fused_bit_exact and (7-224 or 81-160) or test_speculative_path_and_exact_fallback_under_drift'
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  if [ $tool = memcheck ]; then extra="--leak-check no"; fi
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "$SEL" > gpurun_out/sanitizer_${tool}_head.txt 2>&1
  echo "$tool head: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_head.txt | tr '\n' ' ')"
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_gpu_nms.py tests/test_gpu_fp32.py tests/test_gpu_block.py -q -x -m gpu \
      -k "test_doc_examples or test_upstream_assumption_kats or test_edge_cases or test_pred_conv_fp32_split_vs_float64 and 3-5-7 or test_temporal_tip_cell_fp32 or test_temporal_cell_matches_dedicated_kernel or test_detection_block_vs_oracle and 2-shape0" > gpurun_out/sanitizer_${tool}_nms_conv.txt 2>&1
  echo "$tool nms/conv: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_nms_conv.txt | tr '\n' ' ')"
done
