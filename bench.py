#!/usr/bin/env python
"""bench.py -- column-iterations/s of the GLOM column update on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: ``Glom.forward(img, iters=12)`` at BASELINE
configs[1] per GPU (dim=512 L=6 224/14, batch 32, bf16 tensor-core precision), i.e. 12 Jacobi column
updates of 32x256 columns x 6 levels = 589,824 column-iterations per GPU per step.  N > 1 shards the
batch (configs[2]: 32 images per GPU, no data-path collective) => weak scaling.

Printed (rank 0, ONE JSON line):
  value      whole-job column-iterations/s with the images already resident in HBM, device-timed
             (CUDA events on the launch stream, barrier + synchronize both sides, MAX over ranks)
  e2e        same metric through the public API with HOST buffers: pinned-host images -> H2D,
             forward, D2H of the returned state into pinned host memory, all inside the timed region
  roofline   dominant kernel (grouped GEMM1 + GELU, tcgen05): algorithmic FLOPs per launch / its
             average duration from CUDA events recorded around every launch in the timed region
  cpu_baseline  the CPU oracle (numpy port of the reference algorithm) on this box's host cores, on a
             bounded sample of the same workload (rank 0, N = 1 only)

``--impl reference`` times that CPU port alone (the reference itself is a Python package that cannot
travel to the GPU box; the oracle restates it and is pinned against its golden outputs).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(dim=512, levels=6, image_size=224, patch_size=14)
ITERS = 12
BATCH_PER_GPU = 32
N_PATCH = (CFG["image_size"] // CFG["patch_size"]) ** 2
METRIC = "column-iterations/sec (BxNxLxiters) at dim=512 L=6 224/14"
UNIT = "column-iterations/s"


def flops_per_col_iter(d, L, n):
    """Tensor FLOPs per column-iteration: 16 d^2 (2L-1)/L + 4 n d   (SURVEY 8d)."""
    return 16.0 * d * d * (2 * L - 1) / L + 4.0 * n * d


def bytes_per_iter(d, L, n, B, s_state=2, s_w=2):
    """Algorithmic HBM bytes per iteration."""
    return 2 * B * n * L * d * s_state + (2 * L - 1) * (8 * d * d + 5 * d) * s_w + B * n * d * 2 + n * d * 2


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return dict(hbm_gbs=j["hbm_gbs"], tflops=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    tflops_burst=j["bf16_tflops"], source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------ CPU port
def cpu_port_run(batch, iters, reps):
    """Time the numpy oracle on the host cores.  Returns (col-iters/s, seconds per rep, cores)."""
    import numpy as np
    from oracle import glom_oracle as O
    d, L = CFG["dim"], CFG["levels"]
    params = O.synth_params(d, L, CFG["image_size"], CFG["patch_size"], seed=0)
    img = np.random.default_rng(1).standard_normal((batch, 3, CFG["image_size"], CFG["image_size"])).astype(np.float32)
    O.glom_forward(params, img[:1], patch_size=CFG["patch_size"], iters=1, dtype=np.float32)   # warm BLAS
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.glom_forward(params, img, patch_size=CFG["patch_size"], iters=iters, dtype=np.float32)
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return batch * N_PATCH * L * iters / sec, sec, os.cpu_count()


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    # each step is a bounded sample of the same shapes, sized so that warmup + K steps end within a few minutes
    # on any host (the port runs ~5-20 k column-iterations/s): 2 images x 4 iterations = 12,288 column-iterations
    b, it = 2, 4
    for _ in range(min(args.warmup, 1)):
        cpu_port_run(1, 1, 1)
    v, sec, cores = cpu_port_run(b, it, max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1] shapes: dim=512 L=6 224/14; CPU sample batch={b} iters={it} per step",
                   "batch": b, "iters": it},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
                         "sample": f"numpy oracle (BLAS + threaded erf on {cores} cores), batch={b} "
                                   f"iters={it}, median of {max(1, args.steps)} reps, {sec:.2f} s each"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons DURING the timed region: an NVML polling thread (every
    ~5 ms); falls back to `nvidia-smi -lms` if pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.thread = None
        self.samples = []
        self._stop = False
        self.max_mhz = None

    def _poll(self):
        import pynvml
        h = self.handle
        while not self._stop:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                    pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            import pynvml
            self._stop = True
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            bits = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            pmax = max(p for _, p, _ in self.samples)
            load = [s for s in self.samples if s[1] > 0.5 * pmax] or self.samples
            reasons = sorted(k for k, b in bits.items() if any(s[2] & b for s in load))
            return {"sm_mhz": statistics.median(s[0] for s in load), "sm_max_mhz": self.max_mhz,
                    "power_w_max": pmax, "power_w_median_under_load": statistics.median(s[1] for s in load),
                    "samples": len(self.samples), "samples_under_load": len(load), "reasons": reasons,
                    "how": "NVML polled every ~5 ms during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons), "how": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------ ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--iters", type=int, default=ITERS)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train", action="store_true",
                    help="also time a training step (forward with return_all + backward of a loss on all_levels[7,:,:,-1], "
                         "README.md:58-90) and add it to the JSON line as \"train\"")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from glom_pytorch_b200.build import build_library, is_stale
    if rank == 0 and is_stale():
        build_library()
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        dist.barrier(device_ids=[local_rank])
    import glom_pytorch_b200 as G
    from glom_pytorch_b200 import _native
    from glom_pytorch_b200.sharding import shard_range

    if args.warmup < 3:
        args.warmup = 3          # timing rule: at least 3 warm-up steps
    B, T = args.batch_per_gpu, args.iters
    d, L = CFG["dim"], CFG["levels"]
    global_batch = B * world
    s, e = shard_range(global_batch, rank, world)
    assert e - s == B

    torch.manual_seed(0)                                    # identical default init on every rank ...
    model = G.Glom(**CFG, precision=args.precision).to(dev).eval()
    if distributed:                                         # ... and rank 0's weights broadcast over NCCL anyway
        for prm in model.parameters():
            dist.broadcast(prm.data, src=0)

    # synthetic images: 4 rotating pinned host buffers (global batch generated per seed, this rank's shard)
    NBUF = 4
    host_imgs, dev_imgs = [], []
    for i in range(NBUF):
        g = torch.Generator().manual_seed(1 + i)
        full = torch.randn(global_batch, 3, CFG["image_size"], CFG["image_size"], generator=g)
        host_imgs.append(full[s:e].contiguous().pin_memory())
        dev_imgs.append(host_imgs[-1].to(dev))
    host_out = torch.empty(B, N_PATCH, L, d, dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if distributed:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    launches = 0
    with torch.no_grad():
        # -------- device-resident throughput (value) + per-kernel events (roofline), same timed region
        for i in range(args.warmup):
            model(dev_imgs[i % NBUF], iters=T)
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        _native.profile_begin()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for i in range(args.steps):
            model(dev_imgs[i % NBUF], iters=T)
            launches += model.last_launches
        ev1.record(stream)
        barrier()
        ms_dev = ev0.elapsed_time(ev1)
        prof = _native.profile_end()
        clocks = sampler.stop() if rank == 0 else None

        # -------- end to end through the public API with host buffers: every step copies its images from pinned
        # host memory and its result back to pinned host memory inside the timed region.  Copies run on two extra
        # streams (one per direction), double-buffered, so step i's D2H and step i+2's H2D overlap step i+1's compute.
        h2d_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)   # one per copy engine / PCIe direction
        d2h_stream2 = torch.cuda.Stream(dev)           # the 100 MB result goes back as two halves on two DMA queues
        host_outs = [host_out, torch.empty_like(host_out).pin_memory()]

        def e2e_loop(nsteps):
            staged = None
            ready = torch.cuda.Event()
            with torch.cuda.stream(h2d_stream):
                staged = host_imgs[0].to(dev, non_blocking=True)
                ready.record(h2d_stream)
            for i in range(nsteps):
                stream.wait_event(ready)                       # this step's images are on the device
                x = staged
                if i + 1 < nsteps:                             # prefetch the next step's images
                    nxt_ready = torch.cuda.Event()
                    with torch.cuda.stream(h2d_stream):
                        staged = host_imgs[(i + 1) % NBUF].to(dev, non_blocking=True)
                        nxt_ready.record(h2d_stream)
                out = model(x, iters=T)                        # public API call on the compute stream
                done = torch.cuda.Event()
                done.record(stream)
                x.record_stream(stream)
                half = (out.shape[0] + 1) // 2
                for q, (lo, hi) in ((d2h_stream, (0, half)), (d2h_stream2, (half, out.shape[0]))):
                    if lo >= hi:
                        continue
                    with torch.cuda.stream(q):                 # result back to the host
                        q.wait_event(done)
                        host_outs[i % 2][lo:hi].copy_(out[lo:hi], non_blocking=True)
                        out.record_stream(q)
                if i + 1 < nsteps:
                    ready = nxt_ready
            stream.wait_stream(d2h_stream)                     # the last D2H is inside the timed region
            stream.wait_stream(d2h_stream2)

        e2e_loop(max(4, args.warmup))                   # allocator and copy queues reach their steady state
        barrier()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record(stream)
        e2e_loop(args.steps)
        ee1.record(stream)
        barrier()
        ms_e2e = ee0.elapsed_time(ee1)

    train = None
    if args.train:
        model.train()
        tt = 7 if T >= 7 else T
        for _ in range(2):
            model.zero_grad(set_to_none=True)
            model(dev_imgs[0], iters=T, return_all=True)[tt, :, :, -1].square().mean().backward()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        nrep = max(3, args.steps // 4)
        f_ms = b_ms = 0.0
        for i in range(nrep):
            model.zero_grad(set_to_none=True)
            ev[0].record(stream)
            loss = model(dev_imgs[i % NBUF], iters=T, return_all=True)[tt, :, :, -1].square().mean()
            ev[1].record(stream)
            loss.backward()
            ev[2].record(stream)
            torch.cuda.synchronize(dev)
            f_ms += ev[0].elapsed_time(ev[1]); b_ms += ev[1].elapsed_time(ev[2])
        train = {"forward_ms": f_ms / nrep, "backward_ms": b_ms / nrep, "reps": nrep,
                 "value": B * N_PATCH * L * T / ((f_ms + b_ms) / nrep * 1e-3), "unit": "column-iterations/s per GPU (fwd+bwd)",
                 "loss": f"mean(all_levels[{tt}, :, :, -1] ** 2)", "peak_mem_gib": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        model.eval()

    if distributed:
        t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = t.tolist()

    col_iters_step = global_batch * N_PATCH * L * T
    value = col_iters_step * args.steps / (ms_dev * 1e-3)
    e2e_value = col_iters_step * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = measured_peaks()
        rows = B * N_PATCH
        G_ = 2 * L - 1
        kern = {}
        flops = {"gemm1_gelu": 2.0 * rows * 4 * d * d * G_,
                 "gemm2_combine": 2.0 * rows * d * (8 * d * (L - 1) + 4 * d),
                 "attention": 4.0 * N_PATCH * N_PATCH * d * B * L}
        for k, (ms, cnt) in prof.items():
            if cnt:
                kern[k] = {"launches": cnt, "avg_us": ms / cnt * 1e3, "ms_per_step": ms / args.steps}
                if k in flops and args.precision == "bf16":
                    kern[k]["tflops"] = flops[k] / (ms / cnt * 1e-3) / 1e12
        dom = "gemm1_gelu"
        traffic = None
        try:   # per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f)["kernels"][dom]["dram_bytes"] if args.batch_per_gpu == BATCH_PER_GPU else None
        except (OSError, KeyError, ValueError):
            traffic = None
        roof = {"bound": "tensor", "kernel": "gemm_kernel<0,256> (grouped GEMM1 + bias + exact-erf GELU, tcgen05)",
                "achieved": kern.get(dom, {}).get("tflops"), "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": (kern[dom]["tflops"] / peaks["tflops"]) if kern.get(dom, {}).get("tflops") else None,
                "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read+write)",
                "algorithmic_bytes": rows * d * 2 * G_ + G_ * 4 * d * d * 2 + rows * G_ * 4 * d * 2,
                "peak_source": peaks["source"],
                "flops_per_launch": flops[dom],
                "whole_step": {"tflops": flops_per_col_iter(d, L, N_PATCH) * col_iters_step / world /
                               (ms_dev / args.steps * 1e-3) / 1e12,
                               "hbm_gbs_algorithmic": bytes_per_iter(d, L, N_PATCH, B) * T /
                               (ms_dev / args.steps * 1e-3) / 1e9},
                "kernels": kern}
        roof["whole_step"]["frac_tensor"] = roof["whole_step"]["tflops"] / peaks["tflops"]
        roof["whole_step"]["frac_hbm"] = roof["whole_step"]["hbm_gbs_algorithmic"] / peaks["hbm_gbs"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "images_per_s": global_batch * args.steps / (ms_dev * 1e-3),
            "config": {"workload": f"BASELINE configs[{1 if world == 1 else 2}]: dim=512 L=6 224/14 iters={T} "
                                   f"batch={B}/GPU (global {global_batch}), Glom.forward incl. tokeniser",
                       "global_batch": global_batch, "iters": T, "parallelism": f"dp{world} (batch shards, no collective)",
                       "l2": "per-step working set ~1.1 GB (H 369 MB, state 100 MB fp32 + shadows, weights 46 MB) "
                             "> 126 MB L2; input images rotate over 4 buffers; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host_imgs[0].numel() * 4 * world,
                    "d2h_bytes_per_step": host_out.numel() * 4 * world},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
        }
        if train is not None:
            line["train"] = train
        if world == 1 and not args.no_cpu_baseline:
            v, sec, cores = cpu_port_run(4, 3, 5)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
                                    "sample": f"numpy oracle, same shapes, batch=4 iters=3, median of 5 reps "
                                              f"({sec:.2f} s each); BLAS + threaded erf on {cores} cores"}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier(device_ids=[local_rank])
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
