mkdir -p gpurun_out
O=gpurun_out/r2w.out; : > $O
python - >> $O 2>&1 <<'P'
import torch
e=[torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda._sleep(1000); torch.cuda.synchronize()
e[0].record(); torch.cuda._sleep(2_000_000); e[1].record(); torch.cuda.synchronize()
print("idle probe MHz", 2_000_000/(e[0].elapsed_time(e[1])*1e3))
P
VD_TCONV_PREFETCH=0 TC_N=20 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
VD_TCONV_PREFETCH=0 TC_N=200 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
VD_TCONV_PREFETCH=0 TC_N=200 VD_TCONV_DBG=1 timeout 200 python scripts/tconv_scales.py >> $O 2>&1
cat $O
