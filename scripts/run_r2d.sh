mkdir -p gpurun_out
for f in tests/test_gpu_fp32.py tests/test_gpu_head.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu > gpurun_out/$name.log 2>&1; echo "$f exit $?"
  grep -E 'err |identical|redone|passed|failed|FAILED|Error|timed out' gpurun_out/$name.log | tail -n 60
done
