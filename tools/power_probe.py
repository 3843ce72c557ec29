"""Sustained-load clock / power probe: runs config-2 forwards for a few seconds while polling NVML."""
import os, sys, time, threading, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
torch.manual_seed(0)
m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).cuda().eval()
img = torch.randn(32, 3, 224, 224, device="cuda")
samples = []; stop = False
def poll():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.01)
th = threading.Thread(target=poll, daemon=True); th.start()
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
with torch.no_grad():
    for _ in range(3): m(img, iters=12)
    torch.cuda.synchronize()
    t0 = time.time(); n = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    marks = []
    while time.time() - t0 < secs:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): m(img, iters=12)
        e1.record(); torch.cuda.synchronize()
        marks.append((time.time() - t0, e0.elapsed_time(e1) / 20)); n += 20
stop = True; th.join()
for t, ms in marks[::max(1, len(marks) // 12)]:
    near = [s for s in samples if abs(s[0] - t0 - t) < 0.06]
    clk = statistics.median(s[1] for s in near) if near else -1
    pw = statistics.median(s[2] for s in near) if near else -1
    rs = 0
    for s in near: rs |= s[3]
    print(f"t={t:5.2f}s  {ms:6.3f} ms/step  sm {clk} MHz  power {pw:6.1f} W  reasons 0x{rs:x}")
