"""Randomised shape sweep of the forward (both precisions) and the backward (vs the fp32 engine) against the CPU oracle."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import glom_pytorch_b200 as G
from oracle import glom_oracle as O

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 24
bad = 0
for it in range(N):
    dim = int(rng.choice([64, 128, 192, 256, 320, 512]))
    L = int(rng.integers(2, 6))
    p = int(rng.choice([2, 4]))
    side = int(rng.choice([2, 3, 5, 8, 11, 16, 18, 22, 24, 26, 30]))      # 26, 30: n = 676 / 900 > 576 -> consensus key passes
    if side >= 26: dim = min(dim, 128)
    hh, ww = (side, side) if rng.random() < 0.6 else (int(rng.integers(1, side + 1)), side)
    B = int(rng.integers(1, 5)); T = int(rng.integers(1, 4))
    kw = {}
    if rng.random() < 0.3: kw["consensus_self"] = True
    use_radius = rng.random() < 0.3 and (hh, ww) == (side, side)
    if use_radius: kw["local_consensus_radius"] = float(rng.choice([1, 1.5, 2.5]))
    isz = side * p
    params = O.synth_params(dim, L, isz, p, seed=int(rng.integers(1 << 30)))
    img = rng.standard_normal((B, 3, hh * p, ww * p)).astype(np.float32)
    n = hh * ww
    lv = None if rng.random() < 0.5 else (rng.standard_normal((B, n, L, dim)) * float(rng.choice([1, 5]))).astype(np.float32)
    ra = bool(rng.random() < 0.7)                          # return_all or only the final state (ping-pong addressing, S_0 read in place)
    ref = O.glom_forward(params, img, patch_size=p, iters=T, levels=lv, return_all=True, image_size=isz, dtype=np.float64, **kw)
    if not ra: ref = ref[[0, T]]
    line = f"[{it}] d={dim} L={L} grid={hh}x{ww} (n={n}) B={B} T={T} {kw} lv={'y' if lv is not None else 'n'} ra={int(ra)}:"
    for prec in ("fp32", "bf16"):
        m = G.Glom(dim=dim, levels=L, image_size=isz, patch_size=p, precision=prec, **kw)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
        m = m.cuda().eval()
        try:
            with torch.no_grad():
                out = m(torch.from_numpy(img).cuda(), iters=T, levels=None if lv is None else torch.from_numpy(lv).cuda(),
                        return_all=ra).cpu().numpy()
                if not ra: out = np.stack([ref[0].astype(out.dtype), out])
            torch.cuda.synchronize()
        except Exception as e:
            line += f" {prec} EXC {type(e).__name__}: {str(e)[:80]}"; bad += 1; continue
        scale = max(1.0, np.abs(ref).max())
        err = np.abs(out - ref).max() / scale
        rel = max(np.linalg.norm(out[t] - ref[t]) / max(np.linalg.norm(ref[t]), 1e-30) for t in range(1, len(ref)))
        ok = (err <= 1e-4) if prec == "fp32" else (rel <= 1e-2 and err <= 3e-2)
        line += f" {prec} err {err:.2e} rel {rel:.2e} {'ok' if ok else 'FAIL'}"
        bad += (not ok)
    print(line, flush=True)
print("FAILURES:", bad)
