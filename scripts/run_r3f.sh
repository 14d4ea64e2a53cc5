mkdir -p gpurun_out
O=gpurun_out/r3f.out; : > $O
timeout 300 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "fused_tip or temporal or clip" >> $O 2>&1
timeout 300 python scripts/tfused_scales.py 2>&1 | grep '"dbg": "0"' >> $O
timeout 300 python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3f_vid.json 2> gpurun_out/r3f_vid.err
python -c "
import json
for l in open('gpurun_out/r3f_vid.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('vid', round(d['value']), d['ms_per_step'], r['frac'], r['path_frac'], r['kernel_ms'])
" >> $O 2>&1
python scripts/steady_temporal.py > gpurun_out/steady_temporal_plain.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k 'regex:temporal_head_fused' -s 2 -c 1 -o gpurun_out/prof_r02_tfused python scripts/steady_temporal.py > gpurun_out/ncu_r02_tfused.log 2>&1
cat $O
