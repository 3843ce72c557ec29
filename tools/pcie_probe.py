"""Pinned-memory H2D / D2H bandwidth of this box (one stream vs two), for reading the e2e number of bench.py."""
import torch, time
dev = "cuda:0"
x = torch.empty(32, 256, 6, 512, device=dev)
h = torch.empty_like(x, device="cpu").pin_memory()
def bw(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return x.numel() * 4 * reps / (time.perf_counter() - t0) / 1e9
print(f"D2H one stream : {bw(lambda: h.copy_(x, non_blocking=True)):.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    for q, sl in ((s1, slice(0, 16)), (s2, slice(16, 32))):
        with torch.cuda.stream(q): h[sl].copy_(x[sl], non_blocking=True)
print(f"D2H two streams: {bw(two):.1f} GB/s")
print(f"H2D one stream : {bw(lambda: x.copy_(h, non_blocking=True)):.1f} GB/s")
