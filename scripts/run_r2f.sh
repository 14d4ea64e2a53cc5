mkdir -p gpurun_out
for w in vid416_t5_w64 targets_c285_b128 coco608_b64 vid416_b64; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -c 600 gpurun_out/bench_$w.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['unit'], d['ms_per_step'], {k:d['roofline'][k] for k in ('bound','frac','kernel_ms') }, d['roofline'].get('path_frac'), d['cpu_baseline'], d['e2e']['value'])"
done
bash scripts/sanitize.sh
