mkdir -p gpurun_out
O=gpurun_out/r3g.out; : > $O
timeout 300 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "fused_tip or temporal or clip" >> $O 2>&1
timeout 300 python scripts/tfused_scales.py 2>&1 | grep '"dbg": "0"' >> $O
timeout 200 python scripts/tfused_stamps.py 2 2>&1 | tail -1 >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3g_vid.json 2> gpurun_out/r3g_vid.err
python -c "
import json
for l in open('gpurun_out/r3g_vid.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('vid', round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'])
" >> $O 2>&1
cat $O
