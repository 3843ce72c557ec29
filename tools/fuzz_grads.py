"""Randomised shape sweep of the tensor-core backward (bf16 engine) against the fp32 CUDA-core backward."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
bad = 0
for it in range(N):
    dim = int(rng.choice([256, 512]))
    L = int(rng.integers(2, 5))
    p = 2
    side = int(rng.choice([2, 3, 8, 10, 11, 16, 18, 20, 26, 28]))     # 26, 28: n = 676 / 784 > 576 columns
    if side >= 26: dim = 256
    B = int(rng.integers(1, 4)); T = int(rng.integers(1, 3))
    kw = {}
    if rng.random() < 0.3: kw["consensus_self"] = True
    if rng.random() < 0.3: kw["local_consensus_radius"] = float(rng.choice([1.5, 2.5]))
    isz = side * p; n = side * side
    seed = int(rng.integers(1 << 30))
    ms = {}
    for prec in ("fp32", "bf16"):
        torch.manual_seed(seed)
        ms[prec] = G.Glom(dim=dim, levels=L, image_size=isz, patch_size=p, precision=prec, **kw).cuda()
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, isz, isz, generator=g).cuda()
    ra = bool(rng.random() < 0.5)
    with_lv = bool(rng.random() < 0.5)
    lv = torch.randn(B, n, L, dim, generator=g).cuda() if with_lv else None
    cot = torch.randn(((T + 1,) if ra else ()) + (B, n, L, dim), generator=g).cuda()
    grads = {}
    try:
        for prec, m in ms.items():
            x = img.clone().requires_grad_(True)
            l0 = None if lv is None else lv.clone().requires_grad_(True)
            out = m(x, iters=T, levels=l0, return_all=ra)
            (out * cot).sum().backward()
            grads[prec] = {"img": x.grad, **({"levels": l0.grad} if l0 is not None else {}),
                           **{k: q.grad for k, q in m.named_parameters()}}
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[{it}] d={dim} L={L} n={n} B={B} T={T} {kw} EXC {type(e).__name__}: {str(e)[:100]}"); bad += 1; continue
    worst = ("", 0.0)
    for k, ref in grads["fp32"].items():
        got = grads["bf16"][k]
        if ref is None: continue
        rel = (torch.linalg.norm(got - ref) / torch.linalg.norm(ref).clamp_min(1e-30)).item()
        if not np.isfinite(rel) or rel > worst[1]: worst = (k, rel)
    ok = np.isfinite(worst[1]) and worst[1] <= 3e-2
    bad += (not ok)
    print(f"[{it}] d={dim} L={L} n={n} rows={B*n} B={B} T={T} all={ra} lv={with_lv} {kw}: worst rel {worst[1]:.2e} ({worst[0]}) {'ok' if ok else 'FAIL'}", flush=True)
print("FAILURES:", bad)
