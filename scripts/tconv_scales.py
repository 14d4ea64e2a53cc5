"""Per-scale timing of the temporal tip cell (vd_temporal_conv, cfg 4 shapes: 64 windows of T=5 at 416^2), CUDA events.
VD_TCONV_DBG=1 skips the epilogue (mainloop ceiling), VD_TCONV_CTAS=n caps the grid.  One JSON line per scale."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200
from viddet_b200._lib import load, check, ptr, stream_ptr
dev = torch.device("cuda", 0)
B, T = int(os.environ.get("TC_B", 64)), 5
gen = torch.Generator(device=dev).manual_seed(7)
tot = 0.0
for C, hw in ((1024, 13), (512, 26), (256, 52)):
    cell = viddet_b200.TemporalTipConv(C).initialize(generator=torch.Generator().manual_seed(1))
    xs = [torch.randn((B * T, hw, hw, C), generator=gen, device=dev).to(torch.bfloat16) for _ in range(2)]
    y = torch.empty_like(xs[0])
    def run(i):
        check(load().vd_temporal_conv(ptr(xs[i % 2]), ptr(y), B, T, hw, hw, C, ptr(cell._w_taps), ptr(cell._scale), ptr(cell._shift), 0.1, stream_ptr()))
    for i in range(3): run(i)
    torch.cuda.synchronize()
    n = int(os.environ.get("TC_N", 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): run(i)
    e1.record()
    torch.cuda._sleep(2_000_000)            # clock probe: a spin of 2 M SM cycles right behind the loop -> SM MHz while the loop's power state still holds
    e2 = torch.cuda.Event(enable_timing=True); e2.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    mhz = 2_000_000 / (e1.elapsed_time(e2) * 1e3)
    fl = 2.0 * 13 * hw * hw * C * C * B
    tot += ms
    print(json.dumps({"C": C, "hw": hw, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1), "sm_mhz_after": round(mhz), "dbg": os.environ.get("VD_TCONV_DBG"), "ctas": os.environ.get("VD_TCONV_CTAS")}))
print(json.dumps({"total_ms": round(tot, 4)}))
