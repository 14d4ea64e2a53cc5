"""H2D bandwidth of the e2e input path: torch pin_memory() vs write-combined pinned memory (cudaHostAllocWriteCombined)."""
import ctypes, torch, time
rt = ctypes.CDLL("libcudart.so.12") if True else None
torch.cuda.init()
N = 155058176
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
def wc_tensor(n, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_char * n).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)
def bw(src, label):
    for _ in range(3): dev.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dev.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("%-28s pinned=%s  %.1f GB/s" % (label, src.is_pinned(), 10 * N / (e0.elapsed_time(e1) * 1e-3) / 1e9))
a = torch.empty(N, dtype=torch.uint8).pin_memory(); a.fill_(1)
bw(a, "torch pin_memory")
b = wc_tensor(N, 0); b.fill_(1)
bw(b, "cudaHostAlloc default")
c = wc_tensor(N, 4); c.fill_(1)
bw(c, "cudaHostAlloc write-combined")
