mkdir -p gpurun_out
for lvl in 0 1 7 8 9; do
  VD_DEBUG_SKIP_EPILOGUE=$lvl python bench.py --steps 2048 --no-cpu-baseline > gpurun_out/bench_dbg$lvl.json 2> gpurun_out/bench_dbg$lvl.err; python -c "
import json
d=json.load(open('gpurun_out/bench_dbg$lvl.json')); print('dbg level $lvl:', d['ms_per_step']*1e3, 'us/step')"
done
