"""Stage-by-stage GPU diagnostics (development aid; run on the B200 box via gpurun).
Usage: python tools/diag.py <stage>   stages: fp32 | bf16_stages | bf16_full | timing"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import glom_pytorch_b200 as G  # noqa: E402
from glom_pytorch_b200 import _native  # noqa: E402
from oracle import glom_oracle as O  # noqa: E402

DEV = "cuda:0"


def model(dim, L, isz, p, precision, **kw):
    params = O.synth_params(dim, L, isz, p, seed=0)
    m = G.Glom(dim=dim, levels=L, image_size=isz, patch_size=p, precision=precision, **kw)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    return m.to(DEV).eval(), params


def pattern(err, name):
    """Print where errors concentrate: by row%8, row//32, col%8, col//64."""
    bad = err > (1e-2 * max(1.0, err.max()))
    print(f"  [{name}] bad fraction {bad.mean():.4f}")
    if bad.any():
        r, c = np.nonzero(bad)
        print("   rows%8 hist", np.bincount(r % 8, minlength=8).tolist())
        print("   rows//32 hist", np.bincount(r // 32)[:16].tolist())
        print("   cols%8 hist", np.bincount(c % 8, minlength=8).tolist())
        print("   cols//64 hist", np.bincount(c // 64)[:16].tolist())


def stage_fp32():
    for (dim, L, isz, p, B, T) in [(64, 3, 28, 7, 1, 2), (128, 4, 32, 4, 2, 3)]:
        m, params = model(dim, L, isz, p, "fp32")
        img = np.random.default_rng(7).standard_normal((B, 3, isz, isz)).astype(np.float32)
        ref = O.glom_forward(params, img, patch_size=p, iters=T, return_all=True, dtype=np.float64)
        with torch.no_grad():
            out = m(torch.from_numpy(img).to(DEV), iters=T, return_all=True).cpu().numpy()
        print(f"fp32 dim={dim}: max err {np.abs(out - ref).max():.3e}")


def stage_bf16_stages(dim=128, L=4, isz=32, p=4, B=2):
    m, params = model(dim, L, isz, p, "bf16")
    n, d = (isz // p) ** 2, dim
    img = np.random.default_rng(7).standard_normal((B, 3, isz, isz)).astype(np.float32)
    rng = np.random.default_rng(11)
    lv = (rng.standard_normal((B, n, L, d)) * 3).astype(np.float32)
    with torch.no_grad():
        out = m(torch.from_numpy(img).to(DEV), iters=1, levels=torch.from_numpy(lv).to(DEV)).cpu().numpy()
    torch.cuda.synchronize()
    cfg = m.engine_cfg(n)
    ws = m._workspace
    off, nb = _native.workspace_offset(cfg, B, 1, False, 0)
    m128 = (B * n + 127) // 128
    H = ws[off:off + nb].view(torch.bfloat16).float().reshape(2 * L - 1, m128, 4 * d // 64, 128, 64)
    H = H.permute(1, 3, 0, 2, 4).reshape(m128 * 128, 2 * L - 1, 4 * d)[:B * n].cpu().numpy()   # (row, group, 4d)
    off, nb = _native.workspace_offset(cfg, B, 1, False, 1)
    C = ws[off:off + nb].view(torch.bfloat16).float().reshape(B, n, L, d).cpu().numpy()
    P = {k: v.astype(np.float32) for k, v in params.items()}
    tok = O.tokenize(img, P["image_to_tokens.1.weight"], P["image_to_tokens.1.bias"], p, emulate="bf16")
    pos = P["pos_emb.weight"][:n][None, :, None, :]
    lwi = np.concatenate([tok[:, :, None, :], lv], 2)
    w1bu = P["bottom_up.net.1.weight"].reshape(L, 4 * d, d); b1bu = P["bottom_up.net.1.bias"].reshape(L, 4 * d)
    w1td = P["top_down.net.1.weight"].reshape(L - 1, 4 * d, d); b1td = P["top_down.net.1.bias"].reshape(L - 1, 4 * d)
    for l in range(L):
        a = O.bf16_round(lwi[:, :, l, :].reshape(B * n, d))
        want = O.gelu_erf(a @ O.bf16_round(w1bu[l]).T + b1bu[l])
        err = np.abs(H[:, 2 * l] - want)
        print(f"H bu{l}: max err {err.max():.3e} (|want|max {np.abs(want).max():.2f})")
        if err.max() > 5e-2:
            pattern(err, f"H bu{l}")
            print("   got ", H[:2, 2 * l, :6]); print("   want", want[:2, :6])
    for l in range(L - 1):
        a = O.bf16_round((lwi[:, :, l + 2, :] + pos[:, :, 0, :]).reshape(B * n, d))
        want = O.gelu_erf(a @ O.bf16_round(w1td[l]).T + b1td[l])
        err = np.abs(H[:, 2 * l + 1] - want)
        print(f"H td{l}: max err {err.max():.3e}")
        if err.max() > 5e-2:
            pattern(err, f"H td{l}")
    wantC = O.consensus(lv, False, None, emulate="bf16")
    errC = np.abs(C - wantC)
    print(f"C: max err {errC.max():.3e} (|want|max {np.abs(wantC).max():.2f})")
    if errC.max() > 5e-2:
        for l in range(L):
            e = errC[:, :, l, :].reshape(B * n, d)
            print(f"  level {l} max {e.max():.3e}")
            pattern(e, f"C l{l}")
        print("   got ", C[0, :2, 0, :6]); print("   want", wantC[0, :2, 0, :6])
    want = O.glom_forward(params, img, patch_size=p, iters=1, levels=lv, dtype=np.float32, emulate="bf16")
    err = np.abs(out - want)
    print(f"state t+1: max err {err.max():.3e} rel-fro {np.linalg.norm(out - want) / np.linalg.norm(want):.3e}")
    if err.max() > 5e-2:
        for l in range(L):
            e = err[:, :, l, :].reshape(B * n, d)
            print(f"  level {l} max {e.max():.3e}")
            pattern(e, f"S l{l}")


def stage_bf16_full():
    for (dim, L, isz, p, B, T, kw) in [(64, 3, 28, 7, 1, 2, {}), (128, 4, 32, 4, 2, 5, {}),
                                        (128, 4, 32, 4, 2, 3, dict(local_consensus_radius=1.5)),
                                        (256, 3, 64, 4, 2, 2, {}), (512, 6, 224, 14, 2, 3, {})]:
        m, params = model(dim, L, isz, p, "bf16", **kw)
        img = np.random.default_rng(7).standard_normal((B, 3, isz, isz)).astype(np.float32)
        t0 = time.time()
        ref = O.glom_forward(params, img, patch_size=p, iters=T, return_all=True, dtype=np.float32,
                             image_size=isz, **kw)
        with torch.no_grad():
            out = m(torch.from_numpy(img).to(DEV), iters=T, return_all=True).cpu().numpy()
        rel = [float(np.linalg.norm(out[t] - ref[t]) / np.linalg.norm(ref[t])) for t in range(1, T + 1)]
        print(f"bf16 dim={dim} L={L} n={(isz // p) ** 2} {kw}: max err {np.abs(out - ref).max():.3e} rel-fro/step "
              f"{['%.2e' % r for r in rel]} (oracle {time.time() - t0:.1f}s)")


def stage_timing():
    torch.manual_seed(0)
    m = G.Glom(dim=512, levels=6, image_size=224, patch_size=14).to(DEV).eval()
    img = torch.randn(32, 3, 224, 224, device=DEV)
    with torch.no_grad():
        for _ in range(2):
            m(img, iters=12)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m(img, iters=12)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"config2 forward: {ms:.3f} ms/step -> {32 * 256 * 6 * 12 / ms * 1e3:.3e} col-iters/s")
        _native.profile_begin()
        for _ in range(3):
            m(img, iters=12)
        prof = _native.profile_end()
        for k, (t, c) in prof.items():
            if c:
                print(f"  {k}: {t / c * 1e3:.1f} us avg over {c} launches ({t / 3:.3f} ms/step)")
        fl1 = 2 * 8192 * 2048 * 512 * 11; fl2 = 2 * 8192 * 512 * (4096 * 5 + 2048); fla = 4 * 256 * 256 * 512 * 192
        print(f"  gemm1 {fl1 / (prof['gemm1_gelu'][0] / prof['gemm1_gelu'][1] * 1e-3) / 1e12:.1f} TFLOP/s, "
              f"gemm2 {fl2 / (prof['gemm2_combine'][0] / prof['gemm2_combine'][1] * 1e-3) / 1e12:.1f} TFLOP/s, "
              f"attn {fla / (prof['attention'][0] / prof['attention'][1] * 1e-3) / 1e12:.1f} TFLOP/s")


if __name__ == "__main__":
    {"fp32": stage_fp32, "bf16_stages": stage_bf16_stages, "bf16_full": stage_bf16_full,
     "timing": stage_timing}[sys.argv[1]]()
