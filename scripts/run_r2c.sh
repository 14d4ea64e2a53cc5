mkdir -p gpurun_out
for g in 4 2 1; do
python bench.py --steps 20 --warmup 5 --group $g --no-cpu-baseline > gpurun_out/bench20_g$g.json 2> gpurun_out/bench20_g$g.err; echo "bench20 g=$g rc=$?"; tail -c 800 gpurun_out/bench20_g$g.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench20_g$g.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['path_frac'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_events_around_one_eager_launch'], d['roofline']['nms_kernel_ms'], d['details']['speculation'], d['e2e']['value'])
PY
done
for g in 8 4; do
python bench.py --group $g --no-cpu-baseline > gpurun_out/bench2048_g$g.json 2> gpurun_out/bench2048_g$g.err; echo "bench2048 g=$g rc=$?"; tail -c 800 gpurun_out/bench2048_g$g.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench2048_g$g.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['path_frac'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_events_around_one_eager_launch'], d['roofline']['nms_kernel_ms'], d['details']['speculation'])
PY
done
python bench.py --group 4 --pool 12 --no-cpu-baseline > gpurun_out/bench2048_g4_p12.json 2> gpurun_out/bench2048_g4_p12.err; python -c "
import json
d=json.load(open('gpurun_out/bench2048_g4_p12.json')); print('pool12', d['ms_per_step'], d['roofline']['path_frac'])"
python bench.py --group 4 --data same --no-cpu-baseline > gpurun_out/bench2048_g4_same.json 2> gpurun_out/bench2048_g4_same.err; python -c "
import json
d=json.load(open('gpurun_out/bench2048_g4_same.json')); print('same', d['ms_per_step'], d['roofline']['path_frac'])"
python bench.py --group 4 --data video --no-cpu-baseline > gpurun_out/bench2048_g4_video.json 2> gpurun_out/bench2048_g4_video.err; python -c "
import json
d=json.load(open('gpurun_out/bench2048_g4_video.json')); print('video', d['ms_per_step'], d['roofline']['path_frac'], d['details']['speculation'])"
