"""Small-batch forward latency: wall-clock per call (host-bound?) vs device time."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
torch.manual_seed(0)
m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).cuda().eval()
with torch.no_grad():
    for B in (1, 2, 4, 8, 32):
        img = torch.randn(B, 3, 224, 224, device="cuda")
        for _ in range(3): m(img, iters=12)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(20): m(img, iters=12)
        t_launch = (time.perf_counter() - t0) / 20
        e1.record(); torch.cuda.synchronize()
        t_wall = (time.perf_counter() - t0) / 20
        print(f"B={B:3d}: host enqueue {t_launch*1e3:.3f} ms/call, wall {t_wall*1e3:.3f} ms/call, device {e0.elapsed_time(e1)/20:.3f} ms/call")
