"""One config-2 forward (B=32) for profiling under ncu.  argv[1] = iters (default 2)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).cuda().eval()
img = torch.randn(32, 3, 224, 224, device="cuda")
with torch.no_grad():
    out = m(img, iters=iters)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
