import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
from glom_pytorch_b200 import _native
torch.manual_seed(0)
m = G.Glom(dim=1024, levels=8, image_size=384, patch_size=16).cuda().eval()
img = torch.randn(8, 3, 384, 384, device="cuda")
with torch.no_grad():
    for _ in range(2): m(img, iters=16)
    torch.cuda.synchronize()
    _native.profile_begin()
    for _ in range(3): m(img, iters=16)
    torch.cuda.synchronize()
    prof = _native.profile_end()
print(prof)
