# usage: bash scripts/run_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
run() {  # name, extra args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err
  echo "N=$N $name rc=$?"; grep -E "PeerGather|Error|error" gpurun_out/bench_n${N}_$name.err | head -5
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_n${N}_$name.json'))
    print('$name', d['n_gpus'], d['value'], d['unit'], d['ms_per_step'], d['roofline'].get('path_frac'), d['details'].get('sharding'), d['details'].get('gather_verified'), 'e2e', d['e2e']['value'], d['e2e'].get('h2d_only_gbs_per_rank'))
except Exception as e:
    print('no json', e)
PY
}
run voc20 --steps 20 --warmup 5
if [ "$N" = "8" ]; then
  run voc2048
  run vidt5 --workload vid416_t5_w64 --steps 20 --warmup 5
  run targets --workload targets_c285_b128 --steps 20 --warmup 5
  run voc20_nccl --steps 20 --warmup 5 --gather nccl
fi
