mkdir -p gpurun_out
O=gpurun_out/r3j.out; : > $O
timeout 400 python -m pytest tests/test_gpu_head.py -q -x -m gpu -k "fused_tip or temporal or clip" >> $O 2>&1
python - >> $O 2>&1 <<'P'
import torch, time, sys, os
sys.path.insert(0, '.')
import bench, viddet_b200
dev = torch.device("cuda", 0)
W, T, C, size = 64, 5, 30, 416
gen = torch.Generator(device=dev).manual_seed(3)
head = viddet_b200.YOLOV3Head(C, temporal="conv21").initialize(generator=torch.Generator().manual_seed(1234))
head.set_nms(0.45, 400, 100)
big = [torch.empty((W + T - 1, c, size // s, size // s), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last) for c, s in zip(bench.CHANNELS, bench.STRIDES)]
bench.synth_tips(torch, gen, W + T - 1, size, dev, out=big)
for trial in range(2):
    s_ = head.session([viddet_b200.ClipWindows(t, 0, W, T) for t in big])
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); s_.run(); e[1].record(); s_.run(); e[2].record(); s_.run(); e[3].record()
    torch.cuda.synchronize()
    print("cold call (every frame redone by the exact path: failed windows -> tip cells on 16 pairs -> exact head) %.3f ms, then %.3f / %.3f ms; redone frames of the last call %d" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), s_.redone_frames()))
P
cat $O
