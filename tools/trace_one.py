import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
torch.manual_seed(0)
m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).cuda().eval()
img = torch.randn(32, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(2): m(img, iters=2)
    torch.cuda.synchronize()
    os.environ["GLOM_B200_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "128"
    m(img, iters=1)
    torch.cuda.synchronize()
