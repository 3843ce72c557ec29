#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "gradients or denoising or tokenizer or tensor_core_backward" 2>&1 | tail -n 15 ) > gpurun_out/pytest_grad.txt; tail -8 gpurun_out/pytest_grad.txt
timeout 600 python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import glom_pytorch_b200 as G
torch.manual_seed(0)
m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).cuda().train()
img = torch.randn(32, 3, 224, 224, device='cuda', requires_grad=True)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = m(img, iters=12, return_all=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out[7, :, :, -1].square().mean().backward()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: fwd {1e3*(t1-t0):.2f} ms  bwd {1e3*(t2-t1):.2f} ms")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    out = m(img, iters=2, return_all=True); out[1, :, :, -1].square().mean().backward(); torch.cuda.synchronize()
names = sorted({e.key for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, 'self_device_time_total', 0) > 0})
lib = [n for n in names if any(t in n.lower() for t in ('cublas', 'cutlass', 'sgemm', 'gemv', 'cudnn', 'ampere', 'sm90', 'sm100_', 'nvjet'))]
print("library kernels in a training step:", lib)
print("kernels:", [n[:50] for n in names][:60])
PY
