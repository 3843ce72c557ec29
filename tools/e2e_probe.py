"""Where does the end-to-end loop of bench.py lose time?  Device-only vs +D2H vs +H2D vs both (configs[1], B=32)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
torch.manual_seed(0)
dev = torch.device("cuda:0")
m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).to(dev).eval()
himg = [torch.randn(32, 3, 224, 224).pin_memory() for _ in range(4)]
dimg = [h.to(dev) for h in himg]
hout = [torch.empty(32, 256, 6, 512).pin_memory() for _ in range(2)]
stream = torch.cuda.current_stream()
h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
def loop(n, do_h2d, do_d2h):
    staged = dimg[0]; ready = None
    if do_h2d:
        ready = torch.cuda.Event()
        with torch.cuda.stream(h2d):
            staged = himg[0].to(dev, non_blocking=True); ready.record(h2d)
    for i in range(n):
        if do_h2d:
            stream.wait_event(ready); x = staged
            nr = torch.cuda.Event()
            with torch.cuda.stream(h2d):
                staged = himg[(i + 1) % 4].to(dev, non_blocking=True); nr.record(h2d)
        else:
            x = dimg[i % 4]
        out = m(x, iters=12)
        if do_h2d: x.record_stream(stream); ready = nr
        if do_d2h:
            done = torch.cuda.Event(); done.record(stream)
            with torch.cuda.stream(d2h):
                d2h.wait_event(done); hout[i % 2].copy_(out, non_blocking=True); out.record_stream(d2h)
    stream.wait_stream(d2h)
with torch.no_grad():
    for name, a, b in (("device only", 0, 0), ("+D2H", 0, 1), ("+H2D", 1, 0), ("both", 1, 1), ("device only", 0, 0)):
        loop(3, a, b); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); loop(40, a, b); e1.record(); t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"{name:12s}: {e0.elapsed_time(e1) / 40:.3f} ms/step (host enqueue {t_host / 40 * 1e3:.3f} ms/step)  "
              f"reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB")
