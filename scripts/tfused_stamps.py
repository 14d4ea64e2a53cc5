"""Timeline of cluster 0's leader CTA in the fused temporal head kernel (VD_TFUSED_STAMPS=1): per chunk, clock64 of the MMA role
(wait for the free tip accumulator, tip GEMM issue, prediction GEMM issue), the producer (wait for the free prediction-weight buffer)
and epilogue warp 2 (accumulator seen, staged, prediction accumulator seen, decode done).  argv[1] = scale (0 s32, 1 s16, 2 s8)."""
import ctypes, os, sys
import numpy as np, torch
os.environ["VD_TFUSED_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, viddet_b200
from viddet_b200 import _lib
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
W, T, C, size = 64, 5, 30, 416
gen = torch.Generator(device=dev).manual_seed(3)
head = viddet_b200.YOLOV3Head(C, temporal="conv21").initialize(generator=torch.Generator().manual_seed(1234))
head.set_nms(0.45, 400, 100)
big = [torch.empty((W + T - 1, c, size // s, size // s), dtype=torch.bfloat16, device=dev, memory_format=torch.channels_last) for c, s in zip(bench.CHANNELS, bench.STRIDES)]
bench.synth_tips(torch, gen, W + T - 1, size, dev, out=big)
s_ = head.session([viddet_b200.ClipWindows(t, 0, W, T) for t in big])
s_.run(); s_.run(); s_.run()
torch.cuda.synchronize()
os.environ["VD_TFUSED_SCALES"] = str(1 << scale)
off = _lib.load().vd_head_debug_offset(ctypes.byref(s_.params))
s_._ws.view(torch.int64)[off // 8: off // 8 + 16 * 200].zero_()
s_.run(_lib.VD_STAGE_HEAD)
torch.cuda.synchronize()
st = s_._ws.view(torch.int64)[off // 8: off // 8 + 16 * 200].cpu().numpy().reshape(200, 16).astype(np.float64)
used = [i for i in range(200) if st[i, 2] > 0]
t0 = st[used[0], 0]
k = lambda c: (c - t0) / 1000.0 if c > 0 else float("nan")
print("scale %d: %d chunks stamped; times in k-cycles since the first chunk" % (scale, len(used)))
print("chunk | mma: wait_acc  acc_free  tip_issued | pred: at  wp_ok  stg_ok | prod: wp_wait  wp_free | epi: acc_seen  stg_free  staged  pred_seen  decoded")
for i in used[:14] + used[-4:]:
    v = st[i]
    print("%5d | %9.1f %9.1f %10.1f | %8.1f %6.1f %7.1f | %12.1f %8.1f | %12.1f %9.1f %7.1f %10.1f %8.1f" % (
        i, k(v[0]), k(v[1]), k(v[2]), k(v[3]), k(v[4]), k(v[5]), k(v[8]), k(v[9]), k(v[10]), k(v[11]), k(v[12]), k(v[13]), k(v[14])))
u = np.array([st[i] for i in used[2:-2]])
print("per chunk (mean k-cycles): period %.2f | MMA waits for the free accumulator %.2f | tip GEMM issue span %.2f | pred: wait wp %.2f, wait staged %.2f | producer waits for the wp buffer %.2f | epilogue: acc seen -> staged %.2f"
      % (np.diff(u[:, 0]).mean() / 1e3, (u[:, 1] - u[:, 0]).mean() / 1e3, (u[:, 2] - u[:, 1]).mean() / 1e3, (u[:, 4] - u[:, 3]).mean() / 1e3,
         (u[:, 5] - u[:, 4]).mean() / 1e3, (u[:, 9] - u[:, 8]).mean() / 1e3, (u[:, 12] - u[:, 10]).mean() / 1e3))
