"""cuBLAS (torch.bmm, bf16) on the GEMM shapes of K1 / K2 at configs[1], for context next to the fused kernels."""
import torch
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
R, d = 8192, 512
for name, G, M, N, K in (("K1 shape: 10 groups x (8192 x 512) . (512 x 2048)", 10, R, 4 * d, d),
                         ("K2 shape:  6 levels x (8192 x 4096) . (4096 x 512)", 6, R, d, 8 * d),
                         ("square 8192^3", 1, 8192, 8192, 8192)):
    a = torch.randn(G, M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(G, K, N, device="cuda", dtype=torch.bfloat16)
    c = torch.empty(G, M, N, device="cuda", dtype=torch.bfloat16)
    us = t(lambda: torch.bmm(a, b, out=c))
    print(f"{name}: {us:.1f} us, {2 * G * M * N * K / us / 1e6:.0f} TFLOP/s (bare GEMM, bf16 out, no bias / GELU / combine)")
