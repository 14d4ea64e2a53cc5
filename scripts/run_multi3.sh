N=$1
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_$name.json'))
    print('$name', d['n_gpus'], round(d['value']), d['ms_per_step']*1e3, d['roofline']['path_frac'], d['e2e']['value'], d['details'].get('gather_verified'), d['clocks']['reasons'])
except Exception as e:
    print('no json', e)
PY
}
run vidt5_fused --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline
run voc20 --steps 20 --warmup 5 --no-cpu-baseline
