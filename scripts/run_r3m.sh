mkdir -p gpurun_out
timeout 800 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled -k 'regex:temporal_head_fused' -s 30 -c 4 --csv --log-file gpurun_out/r02_ncu_tfused_clip_w64_traffic.csv python bench.py --workload vid416_t5_w64 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_tfused_clip.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_ncu_tfused_clip_w64_traffic.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[0], r[12], r[14], r[13])
P
