mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_block.py -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_head.py -q -m gpu -k "grouped_launch" 2>&1 | tail -3
python scripts/bench_block.py neck 2> gpurun_out/neck.err | tee gpurun_out/r02_neck.jsonl; tail -3 gpurun_out/neck.err
