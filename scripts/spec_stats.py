"""State of the speculative path after a few steady-state calls: threshold, failed frames, emitted counts per frame."""
import sys, os, torch, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viddet_b200, bench
wl = sys.argv[1] if len(sys.argv) > 1 else "voc416_b64"
C, size, frames = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1234)
head = viddet_b200.YOLOV3Head(C).initialize(generator=torch.Generator().manual_seed(1234))
s = head.session(bench.synth_tips(torch, gen, frames, size, dev))
hw = [(size // st) ** 2 for st in bench.STRIDES]
anc = 3 * sum(hw); tif = sum((h + 127) // 128 for h in hw); F = frames
def al(x): return (x + 255) // 256 * 256
n1 = (tif + 7) // 8
off = 256 + 256 + al(F * 4096 * 4) + al(F * anc * 16) + al(F * tif * 1024 * 8) + al(F * tif * 4) * 2 + al(F * 4) + al(F * 256)
off_spec_lists = off; off += al(F * 2048 * 8)
off_cnt = off; off += al(F * 4)
off_state = off
for call in range(6):
    s.run(viddet_b200._lib.VD_STAGE_HEAD)
    torch.cuda.synchronize()
    cnt = s._ws[off_cnt: off_cnt + F * 4].view(torch.int32).cpu()
    st_before = s._ws[off_state: off_state + 20].view(torch.int32).cpu().tolist()
    s.run(viddet_b200._lib.VD_STAGE_NMS)
    torch.cuda.synchronize()
    st = s._ws[off_state: off_state + 20].view(torch.int32).cpu().tolist()
    tau = s._ws[off_state + 256: off_state + 256 + F * 4].view(torch.float32).cpu()
    print("call %d: emitted per frame min %d med %d max %d, failed frames %d, next tau min %.4f med %.4f max %.4f" % (
        call, int(cnt.min()), int(cnt.median()), int(cnt.max()), st[4], float(tau.min()), float(tau.median()), float(tau.max())))
