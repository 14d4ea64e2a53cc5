"""A few YOLODetectionBlockV3 forward passes at the VOC-416 s16 scale (64 frames): target of ncu for conv_bn_lrelu_kernel."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200
g = torch.Generator().manual_seed(0)
blk = viddet_b200.YOLODetectionBlockV3(256, "2", in_channels=768).initialize(generator=g)
x = torch.randn(64, 768, 26, 26, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for _ in range(3): blk(x)
torch.cuda.synchronize(); print("done")
