mkdir -p gpurun_out
O=gpurun_out/r3k.out; : > $O
timeout 400 python -m pytest tests/test_gpu_block.py tests/test_gpu_ref_exec.py -q -x -m gpu >> $O 2>&1
timeout 400 python scripts/bench_block.py 64 1 neck > gpurun_out/r3k_block.jsonl 2>> $O
python - >> $O <<'P'
import json
for l in open('gpurun_out/r3k_block.jsonl'):
    if l.startswith('{'):
        d=json.loads(l)
        if 'neck' in d: print('NECK', d.get('neck')[:60], d.get('ms'), d.get('frames_per_s'), d.get('frac_tensor_peak'))
        else: print(d.get('cell', d.get('what','?')), d.get('ms'), d.get('frac_tensor_peak', d.get('tensor_frac')))
P
cat $O
