"""A few temporal (conv21) fused calls: target of ncu for temporal_conv_kernel."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
windows, T, C, size = 32, 5, 30, 416
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C, temporal="conv21").initialize(generator=torch.Generator().manual_seed(1234))
tips = []
for c, s in zip(bench.CHANNELS, bench.STRIDES):
    h = size // s
    x = torch.randn((windows * T, c, h, h), generator=gen, device=dev)
    x = torch.where(x > 0, x, 0.1 * x).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    tips.append(x.reshape(windows, T, c, h, h))
sess = head.session(tips)
for _ in range(3): sess.run()
torch.cuda.synchronize(); print("done")
