N=$1
mkdir -p gpurun_out
run() {
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_$name.json'))
    print('$name', d['n_gpus'], round(d['value']), d['ms_per_step']*1e3, d['details'].get('timed_region_ms_per_rank'), d['clocks'])
except Exception as e:
    print('no json', e)
PY
}
run voc20a --steps 20 --warmup 5
run voc20b --steps 20 --warmup 5
run voc20_g5 --steps 20 --warmup 5 --group 5
run voc20_w40 --steps 20 --warmup 40
