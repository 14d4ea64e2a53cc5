mkdir -p gpurun_out
for f in tests/test_gpu_fp32.py tests/test_gpu_head.py; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -q -s -m gpu > gpurun_out/$name.log 2>&1; echo "$f exit $?"
  grep -E 'err |identical|redone|passed|failed|FAILED|Error' gpurun_out/$name.log | tail -n 60
done
python scripts/graph_latency.py 2> gpurun_out/gl.err | tee gpurun_out/graph_latency.jsonl
VD_LIB=viddet_b200/variants/libviddet_b200_fb8.so python scripts/graph_latency.py 2>> gpurun_out/gl.err | tee gpurun_out/graph_latency_fb8.jsonl
tail -5 gpurun_out/gl.err
