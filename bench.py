#!/usr/bin/env python
"""bench.py — headline benchmark of the InfoNCE / IIC loss hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg2]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one fwd+bwd of SupConLoss1 over one batch of synthetic, L2-normalised embeddings (BASELINE.json
`metric`: InfoNCE fwd+bwd pairs/s; pairs = N^2 per step).  Workload (all N): BASELINE config 4's shape, global InfoNCE
N=65536, d=256, bf16, meta-labels randint(0,4096) — the shape north_star quotes the tensor-core target on; it fits
one GPU because the N x N matrix is never materialised, and for N>1 the same problem is row-sharded (strong scaling:
all-gather of embeddings + row statistics, no gradient collective).  `--workload cfg2` runs config 2's N=32768.

The JSON line also carries the IIC-seg leg (config 3: 32 x 10 x 224 x 224 fp32, padding 1; pixels/s against the HBM
roofline) under "iic", a CPU baseline ("cpu_baseline": the C/OpenMP oracle port timed on the host cores on a bounded
sample), kernel-level rooflines measured with CUDA events around the C-ABI calls, and SM clocks sampled during the
timed region.  `--impl reference` times the CPU port alone (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg4": dict(N=65536, d=256, classes=4096, name="cfg4 global InfoNCE N=65536 d=256 bf16 meta-labels(4096)"),
    "cfg2": dict(N=32768, d=256, classes=0, name="cfg2 dense InfoNCE N=32768 (16 img x 1024 px x 2 views) d=256 bf16 self-labels"),
    # the rest of the family on cfg4's shape: SelfPacedSupConLoss / exclude_other_pos on the tensor path
    "selfpaced": dict(N=65536, d=256, classes=4096, variant=3, gamma=12.0,
                      name="cfg4 shape, SelfPacedSupConLoss(weight_update='soft', gamma=12) N=65536 d=256 bf16 meta-labels(4096)"),
    "selfpaced_hard": dict(N=65536, d=256, classes=4096, variant=2, gamma=11.0,
                           name="cfg4 shape, SelfPacedSupConLoss(weight_update='hard', gamma=11) N=65536 d=256 bf16 meta-labels(4096)"),
    "exclude": dict(N=65536, d=256, classes=4096, variant=1,
                    name="cfg4 shape, SupConLoss1(exclude_other_pos=True) N=65536 d=256 bf16 meta-labels(4096)"),
    "d128": dict(N=65536, d=128, classes=4096, name="cfg4 shape with d=128: N=65536 bf16 meta-labels(4096)"),
    # fp32 embeddings under torch.autocast (the reference's default AMP config hands fp32 to the criterion and runs the GEMM in
    # half precision): the module follows autocast and takes the tensor kernels
    # fp32 embeddings outside autocast: [hi | lo] bf16 split on the tensor kernels (three MMA terms per product, fp32 parity)
    "fp32": dict(N=65536, d=256, classes=4096, dtype="f32",
                 name="cfg4 shape, fp32 inputs (bf16 hi/lo split, 3 MMA terms): N=65536 d=256 meta-labels(4096)"),
    "amp": dict(N=65536, d=256, classes=4096, dtype="f32", autocast=True,
                name="cfg4 shape, fp32 inputs under torch.autocast(bf16): N=65536 d=256 meta-labels(4096)"),
}
IIC_CFG = dict(B=32, K=10, H=224, W=224, pad=1)
CPU_SAMPLE_N = 8192          # the reference materialises ~13 N x N fp32 tensors: 8192 is what fits / finishes in seconds


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """SM clock + throttle reasons sampled every ~10 ms DURING the timed region (NVML in a thread; nvidia-smi fallback)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons = index, [], None, set()
        self._stop = threading.Event()
        self.thread = None

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES so that LOCAL_RANK maps to the right physical GPU
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
        except Exception:  # noqa
            self._smi()

    def _smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                     "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        for line in proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                try:
                    self.sm.append(float(parts[0])); self.mx = float(parts[1])
                except ValueError:
                    pass
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            if self._stop.is_set():
                proc.terminate()
                break

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def ncu_traffic():
    """DRAM bytes per launch measured by ncu (profiles/r2_traffic.json, exported from the committed --set full captures)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _timed(fn, steps, warmup):
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fn()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return statistics.median(times), times


def cpu_supcon_baseline(steps=3, warmup=1, seed=0, want_port=True):
    """SupConLoss1 fwd+bwd on the host cores, bounded sample N=8192, d=256, fp32 (the reference materialises ~13 N x N fp32
    tensors: 8192 is what fits).  Primary figure: the reference's OWN module (baseline/_ref, torch CPU) when it is staged
    (kind "reference"); the C/OpenMP port (oracle/oracle.c) is timed next to it (kind "port" when it is all there is).
    Thread counts are set explicitly: torchrun exports OMP_NUM_THREADS=1 to its ranks."""
    import numpy as np
    from oracle import c_oracle, reference_arm
    cores = host_cores()
    c_oracle.set_num_threads(cores)
    rng = np.random.default_rng(seed)
    N, d = CPU_SAMPLE_N, 256
    z = rng.standard_normal((N, d), dtype=np.float32)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    lab_half = rng.integers(0, 512, N // 2).astype(np.int32)
    lab = np.tile(lab_half, 2)
    port = None
    if want_port or not reference_arm.available():
        t_port, _ = _timed(lambda: c_oracle.supcon_fwd_bwd(z, lab, t=0.07, prec=0), steps, warmup)
        port = dict(value=N * N / t_port, unit="pairs/s", cores=c_oracle.num_threads(), kind="port", s_per_step=t_port,
                    sample=f"SupConLoss1 fwd+bwd, N={N}, d={d}, fp32, C/OpenMP port of contrastive.py (oracle/oracle.c), "
                           f"median of {steps} after {warmup} warm-up")
    if not reference_arm.available():
        return port, port["s_per_step"]
    import torch
    torch.set_num_threads(cores)
    SupCon, _, _ = reference_arm.load()
    zt = torch.from_numpy(z)
    target = lab_half.tolist()
    crit = SupCon()

    def ref_step():
        f1 = zt[:N // 2].clone().requires_grad_()
        f2 = zt[N // 2:].clone().requires_grad_()
        crit(f1, f2, target=target).backward()
    t_ref, _ = _timed(ref_step, steps, warmup)
    base = dict(value=N * N / t_ref, unit="pairs/s", cores=torch.get_num_threads(), kind="reference", s_per_step=t_ref,
                sample=f"the reference's own contrastyou.losses.contrastive.SupConLoss1 (baseline/_ref, torch {torch.__version__} "
                       f"CPU, fp32) fwd+bwd, N={N}, d={d}, median of {steps} after {warmup} warm-up")
    if port is not None:
        base["port"] = port
    return base, t_ref


def cpu_iic_baseline(steps=3, warmup=1, seed=0):
    """IIDSegmentationLoss(padding=1) fwd+bwd at config 3's FULL batch (32 x 10 x 224 x 224 fp32) on the host cores"""
    import numpy as np
    from oracle import c_oracle, reference_arm
    cores = host_cores()
    c_oracle.set_num_threads(cores)
    rng = np.random.default_rng(seed)
    B, K, H, W, pad = (IIC_CFG[k] for k in ("B", "K", "H", "W", "pad"))

    def sm(a):
        e = np.exp(a - a.max(1, keepdims=True))
        return (e / e.sum(1, keepdims=True)).astype(np.float32)
    x = sm(2 * rng.standard_normal((B, K, H, W), dtype=np.float32))
    y = sm(2 * rng.standard_normal((B, K, H, W), dtype=np.float32))
    t_port, _ = _timed(lambda: c_oracle.iic_fwd_bwd(x, y, pad, prec=0), steps, warmup)
    port = dict(value=B * H * W / t_port, unit="pixels/s", cores=c_oracle.num_threads(), kind="port", s_per_step=t_port,
                sample=f"IIDSegmentationLoss(padding=1) fwd+bwd, {B}x{K}x{H}x{W} fp32 (config 3, full batch), C/OpenMP port of "
                       f"discreteMI.py, median of {steps} after {warmup} warm-up")
    if not reference_arm.available():
        return port
    import torch
    torch.set_num_threads(cores)
    _, _, IIC = reference_arm.load()
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    crit = IIC(padding=pad)

    def ref_step():
        a, b = xt.clone().requires_grad_(), yt.clone().requires_grad_()
        crit(a, b).backward()
    t_ref, _ = _timed(ref_step, steps, warmup)
    return dict(value=B * H * W / t_ref, unit="pixels/s", cores=torch.get_num_threads(), kind="reference", s_per_step=t_ref,
                sample=f"the reference's own contrastyou.losses.discreteMI.IIDSegmentationLoss(padding=1) (baseline/_ref, torch CPU, "
                       f"fp32) fwd+bwd, {B}x{K}x{H}x{W} (config 3, full batch), median of {steps} after {warmup} warm-up",
                port=port)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (baseline/_ref: its unmodified SupConLoss1 on
    torch CPU; the C/OpenMP port when that directory is absent) on ALL host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = 1 if args.warmup else 0
    base, t = cpu_supcon_baseline(steps=steps, warmup=warm)
    wl = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "InfoNCE fwd+bwd pairs/s", "value": base["value"], "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "host_cores": host_cores(),
                   "note": f"CPU reference timed on a bounded sample N={CPU_SAMPLE_N} (it materialises ~13 N x N fp32 tensors); "
                           "pairs/s is flat in N on CPU"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU legs
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-iic", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="additionally measure the step replayed from CUDA graphs (deferred_checks modules; extra key 'graphed')")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from contrast_you_b200 import _lib as L
    from contrast_you_b200.losses import SupConLoss1, SelfPacedSupConLoss, IIDSegmentationLoss
    from contrast_you_b200.losses.contrastive import _canonical_labels, sort_rows_by_label
    from contrast_you_b200 import distributed as cyd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    lib = L.load()
    W_ = max(3, args.warmup)
    K_ = max(1, args.steps)
    wl = WORKLOADS[args.workload]
    N, d = wl["N"], wl["d"]
    n = N // 2
    n_loc = n // world
    assert n % world == 0

    # ---- synthetic inputs (seeded; identical on every rank, each rank keeps its slice) in PINNED HOST memory
    gen = torch.Generator().manual_seed(0)
    in_dtype = torch.float32 if wl.get("dtype") == "f32" else torch.bfloat16
    z_host = torch.nn.functional.normalize(torch.randn(N, d, generator=gen), dim=1).to(in_dtype)
    variant, gamma = int(wl.get("variant", 0)), float(wl.get("gamma", 1e6))
    lab_host = (torch.randint(0, wl["classes"], (n,), generator=gen) if wl["classes"] else torch.arange(n)).to(torch.int32)
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    f1_host = z_host[:n][sl].contiguous().pin_memory()
    f2_host = z_host[n:][sl].contiguous().pin_memory()
    lab_loc_host = lab_host[sl].contiguous().pin_memory()
    f1_dev, f2_dev, lab_dev = f1_host.to(dev), f2_host.to(dev), lab_loc_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def make_criterion(path, sharded):
        if variant in (2, 3):
            assert not sharded, "the self-paced workloads are single-GPU bench lines"
            c = SelfPacedSupConLoss(weight_update="hard" if variant == 2 else "soft", path=path)
            c.set_gamma(gamma)
            return c
        if sharded:
            return cyd.ShardedSupConLoss(exclude_other_pos=variant == 1, group=None, path=path)
        return SupConLoss1(exclude_other_pos=variant == 1, path=path)
    crit = make_criterion(args.path, world > 1)
    use_autocast = bool(wl.get("autocast"))

    def step(a, b, lab, criterion=None):
        a = a.detach().requires_grad_()
        b = b.detach().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_autocast):
            loss = (criterion or crit)(a, b, target=lab)
        loss.backward()
        return loss, a.grad, b.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    counted = {"launches": 0}

    def timed_loop(fn, warmup, steps, gate_ms=0.0):
        """per-step CUDA-event times, summed.  gate_ms > 0: the stream first spins for about that long (torch.cuda._sleep) so that
        the host has the steps queued before the device starts on them — the events then bracket device work only, as they do in
        a training loop whose GPU queue is deep; without it a step shorter than its own enqueue time is timed at the host's pace."""
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        launches0 = lib.cy_launch_count()
        if gate_ms > 0:
            torch.cuda._sleep(int(gate_ms * 2.0e6))             # ~2 GHz: at least gate_ms
        for _ in range(steps):
            flush.zero_()                                   # L2 flush between timed iterations (outside the brackets)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        barrier()
        counted["launches"] = lib.cy_launch_count() - launches0      # kernels of libcontrastyou_b200.so in the timed region
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

    # ---- device-resident leg (`value`)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    total_ms = timed_loop(lambda: step(f1_dev, f2_dev, lab_dev), W_, K_)
    clocks = sampler.stop() if sampler else None
    ms_per_step = total_ms / K_
    value = N * N / (ms_per_step * 1e-3)
    value_launches = counted["launches"]

    # ---- parity of exactly what was timed, OUTSIDE the timed region (VERDICT r1 #1b): N=1 against the float C oracle at
    # full size; N>1: the sharded loss and this rank's gradient rows against a single-process evaluation of the gathered
    # problem on the same GPU (max over ranks)
    def parity_block():
        loss, ga, gb = step(f1_dev, f2_dev, lab_dev)
        got_loss = float(loss.item())
        if world == 1 and (variant != 0 or use_autocast or d != 256) and not (in_dtype == torch.float32 and not use_autocast):
            # the C oracle restates SupConLoss1's default variant; the other members of the family are checked at full size
            # against this library's fp32 CUDA-core path, which the -m gpu suite pins to the reference's fixtures
            ref, ra, rb_ = step(f1_dev, f2_dev, lab_dev, make_criterion("simt", False))
            ref_loss, ref_grad = float(ref.item()), torch.cat([ra, rb_]).float().cpu()
            got = torch.cat([ga, gb]).float().cpu()
            against = f"fp32 CUDA-core path of this library (pinned to the reference fixtures by the -m gpu suite) at the full N={N}"
        elif world == 1:
            if args.no_cpu:
                return None
            from oracle import c_oracle
            c_oracle.set_num_threads(host_cores())
            t0 = time.perf_counter()
            import numpy as np
            o = c_oracle.supcon_fwd_bwd(torch.cat([f1_dev, f2_dev]).float().cpu().numpy(),
                                        np.tile(lab_dev.cpu().numpy().astype(np.int32), 2), t=0.07, prec=0)
            ref_loss, ref_grad = o["loss"], torch.from_numpy(o["grad"])
            got = torch.cat([ga, gb]).float().cpu()
            against = f"C oracle (oracle/oracle.c, float32 dots, float64 row sums) at the full N={N}, {time.perf_counter() - t0:.1f} s"
        else:
            def gather(t):
                out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                dist.all_gather_into_tensor(out, t.contiguous())
                return out
            a_all, b_all, l_all = gather(f1_dev), gather(f2_dev), gather(lab_dev)
            a_all.requires_grad_(); b_all.requires_grad_()
            single = SupConLoss1(path=args.path)
            ref = single(a_all, b_all, target=l_all)
            ref.backward()
            ref_loss = float(ref.item())
            ref_grad = torch.cat([a_all.grad[sl], b_all.grad[sl]]).float().cpu()
            got = torch.cat([ga, gb]).float().cpu()
            against = f"single-process SupConLoss1 on the gathered N={N} problem, evaluated on every rank"
        loss_rel = abs(got_loss - ref_loss) / abs(ref_loss)
        grad_rel = float((got - ref_grad).abs().max() / ref_grad.abs().max())
        if world > 1:
            t = torch.tensor([loss_rel, grad_rel], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            loss_rel, grad_rel = float(t[0]), float(t[1])
        # hard self-paced weights are a step function of -log p: pairs within float rounding of gamma flip between the two
        # evaluations, each flip moving one weight by 1/c_i — the gradient bar is looser there, the loss bar is not
        gtol = 5e-2 if variant == 2 else (1e-4 if (in_dtype == torch.float32 and not use_autocast) else 1e-2)
        return {"loss": got_loss, "ref_loss": ref_loss, "loss_rel": loss_rel, "grad_rel": grad_rel, "against": against,
                "tolerance": {"loss_rel": 1e-4, "grad_rel": gtol}, "ok": bool(loss_rel <= 1e-4 and grad_rel <= gtol)}
    parity = parity_block()

    # ---- end-to-end leg: host buffers -> module -> loss back on the host, every step.  `serial`: copy, compute, read, one after
    # the other; headline: the same with contrast_you_b200.prefetch.HostPrefetcher (step k+1's inputs are copied on a copy stream
    # while step k computes — one H2D copy of the step's inputs per step in both forms, all inside the timed brackets)
    def e2e_step():
        a = f1_host.to(dev, non_blocking=True)
        b = f2_host.to(dev, non_blocking=True)
        lab = lab_loc_host.to(dev, non_blocking=True)
        loss, _, _ = step(a, b, lab)
        return loss.item()                                   # D2H read of the step's result
    e2e_serial_ms = timed_loop(e2e_step, 3, K_) / K_
    from contrast_you_b200.prefetch import HostPrefetcher
    pf = HostPrefetcher([f1_host, f2_host, lab_loc_host], dev)

    def e2e_step_pipelined():
        a, b, lab = pf.next()
        loss, _, _ = step(a, b, lab)
        return loss.item()
    e2e_ms = timed_loop(e2e_step_pipelined, 3, K_) / K_
    del pf
    h2d = f1_host.numel() * f1_host.element_size() * 2 + lab_loc_host.numel() * 4
    e2e = {"value": N * N / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": e2e_ms, "serial_ms_per_step": e2e_serial_ms,
           "pipeline": "double-buffered H2D on a copy stream (HostPrefetcher): step k+1's inputs are copied while step k computes"}

    # ---- the north_star's contract-named backward, design (i): strip contribution to all N rows + reduce-scatter of dZ, timed
    # beside the default design (ii) (complete gradient of the owned rows from the gathered row statistics)
    rs_variant = None
    if world > 1 and variant == 0:
        rs_crit = cyd.ShardedSupConLoss(group=None, path=args.path, backward_design="reduce_scatter")
        rs_ms = timed_loop(lambda: step(f1_dev, f2_dev, lab_dev, rs_crit), 1, 3) / 3
        rs_variant = {"ms_per_step": rs_ms, "value": N * N / (rs_ms * 1e-3), "unit": "pairs/s",
                      "bytes_reduce_scattered_per_rank": N * d * 4,
                      "note": "design (i): every rank forms dL/dS for its row strip, contributes G Z to its rows and G^T Z to all N "
                              "columns (fp32, stock torch ops on the strip) and the [N, d] gradient is reduce-scattered (NCCL); "
                              "same numbers as the default design (ii), which needs no gradient collective"}

    # ---- optional: the same step replayed from CUDA graphs (forward and backward captured once, collectives included)
    graphed = None
    if args.graph:
        gcrit = (SupConLoss1(path=args.path, deferred_checks=True) if world == 1
                 else cyd.ShardedSupConLoss(group=None, path=args.path, deferred_checks=True))
        sa, sb, sl_ = f1_dev.clone().requires_grad_(), f2_dev.clone().requires_grad_(), lab_dev.clone()
        gfn = torch.cuda.make_graphed_callables(lambda a, b, lab: gcrit(a, b, target=lab), (sa, sb, sl_))

        def gstep(a, b, lab):
            sa.detach().copy_(a, non_blocking=True)
            sb.detach().copy_(b, non_blocking=True)
            sl_.copy_(lab, non_blocking=True)
            sa.grad = None
            sb.grad = None
            loss = gfn(sa, sb, sl_)
            loss.backward()
            return loss

        g_ms = timed_loop(lambda: gstep(f1_dev, f2_dev, lab_dev), W_, K_) / K_
        ge_ms = timed_loop(lambda: gstep(f1_host, f2_host, lab_loc_host).item(), 3, K_) / K_
        gcrit.raise_if_flagged()
        graphed = {"value": N * N / (g_ms * 1e-3), "ms_per_step": g_ms,
                   "e2e": {"value": N * N / (ge_ms * 1e-3), "ms_per_step": ge_ms},
                   "note": "forward + backward replayed from CUDA graphs (torch.cuda.make_graphed_callables); NaN / normalisation "
                           "checks accumulate in device counters read once after the loop; inputs are copied into the graphs' "
                           "static buffers inside the timed region"}

    # ---- kernel-level roofline: CUDA events around the C-ABI calls themselves (single GPU problem: this rank's rows)
    # same row order as the modules use: each rank's row block sorted by label
    with torch.no_grad():
        if world == 1:
            labels = _canonical_labels(lab_dev, n, dev)
            z_all, labels = sort_rows_by_label(torch.cat([f1_dev, f2_dev]), labels)
            z_all = z_all.contiguous()
            rb, re = 0, N
        else:
            lab_loc = _canonical_labels(lab_dev, n_loc, dev)
            order = torch.argsort(lab_loc)
            z_all = cyd.gather_rank_major(torch.cat([f1_dev, f2_dev]).index_select(0, order).contiguous())
            labels = cyd.gather_rank_major(lab_loc.index_select(0, order).contiguous())
            rb, re = cyd.row_range(n_loc)
    path = {"auto": 0, "simt": 1, "tcgen05": 2}[args.path]
    stats = torch.empty(L.CY_NSTAT, N, dtype=torch.float32, device=dev)
    xstat = torch.zeros(N, 4, dtype=torch.float32, device=dev)
    out4 = torch.zeros(8, device=dev)
    ws_b = lib.cy_infonce_workspace_bytes(N, d, L.CY_F32_SPLIT if (in_dtype == torch.float32 and not use_autocast) else L.CY_BF16,
                                          variant, path)
    ws = torch.empty(ws_b, dtype=torch.uint8, device=dev)
    one = torch.ones(1, device=dev)
    dz = torch.empty_like(z_all)
    st = L.stream_ptr()

    kdt, kld = L.CY_BF16, d   # the kernels see what the module hands them: bf16 (fp32 inputs under autocast are cast by the module)
    if in_dtype != torch.bfloat16 and use_autocast:
        z_all = z_all.to(torch.bfloat16)
        dz = torch.empty_like(z_all)
    elif in_dtype != torch.bfloat16:      # fp32 outside autocast: [hi | lo] bf16 halves, fp32 gradient
        assert world == 1
        zs = torch.empty(N, 2 * d, dtype=torch.bfloat16, device=dev)
        L.check(lib.cy_infonce_pack_split(z_all[:n].data_ptr(), z_all[n:].data_ptr(), n, d, d, d, None, zs.data_ptr(), None, st), "split")
        dz = torch.empty(N, d, dtype=torch.float32, device=dev)
        z_all, kdt, kld = zs, L.CY_F32_SPLIT, 2 * d

    def k_fwd():
        L.check(lib.cy_infonce_fwd(z_all.data_ptr(), kdt, N, d, kld, labels.data_ptr(), None, rb, re, 1 / 0.07, variant, path,
                                   stats.data_ptr(), xstat.data_ptr(), ws.data_ptr(), ws_b, st), "fwd")

    def k_fwd2():
        L.check(lib.cy_infonce_fwd_pass2(z_all.data_ptr(), kdt, N, d, kld, labels.data_ptr(), None, rb, re, 1 / 0.07, variant, gamma,
                                         path, stats.data_ptr(), xstat.data_ptr(), ws.data_ptr(), ws_b, st), "fwd2")

    def k_fin():
        if variant != 0:
            k_fwd2()
        if world > 1:
            cyd.gather_rows_(xstat)
        L.check(lib.cy_infonce_loss(N, variant, xstat.data_ptr(), out4.data_ptr(), None, None, ws.data_ptr(), ws_b, st), "loss")

    def k_bwd():
        L.check(lib.cy_infonce_bwd(z_all.data_ptr(), kdt, N, d, kld, labels.data_ptr(), None, rb, re, 1 / 0.07, variant, gamma, path,
                                   xstat.data_ptr(), one.data_ptr(), dz.data_ptr(), d, ws.data_ptr(), ws_b, st), "bwd")
    k_fwd(); k_fin()
    kreps = max(3, min(K_, 10))
    fwd_ms = timed_loop(k_fwd, 2, kreps) / kreps
    fwd2_ms = None
    if variant != 0:        # second sweep of exclude / self-paced: only the column tiles that can hold positives
        k_fwd()
        fwd2_ms = timed_loop(k_fwd2, 2, kreps) / kreps
        k_fwd(); k_fin()
    bwd_ms = timed_loop(k_bwd, 2, kreps) / kreps
    rows = re - rb
    fwd_tf = 2.0 * rows * N * d / (fwd_ms * 1e-3) / 1e12
    bwd_tf = 4.0 * rows * N * d / (bwd_ms * 1e-3) / 1e12
    peak_tf = peaks["bf16_tflops"]
    traffic = ncu_traffic()
    roofline = {"bound": "tensor", "kernel": "cy_infonce_bwd (recompute S tile + W.Z, 4*rows*N*d FLOP per launch)",
                "achieved": bwd_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": bwd_tf / peak_tf,
                # DRAM bytes per launch from the committed ncu capture of this exact workload (1 GPU, N=65536)
                "traffic": traffic.get("infonce_bwd_tc_kernel<256,0,0,0> N=65536 d=256 bf16", {}).get("bytes")
                if (world == 1 and N == 65536 and path != 1) else None,
                "peak_source": peaks["source"] + " bf16 burst",
                "fwd": {"kernel": "cy_infonce_fwd (2*rows*N*d FLOP)", "ms": fwd_ms, "achieved": fwd_tf, "frac": fwd_tf / peak_tf},
                "bwd_ms": bwd_ms, "fwd_pass2_ms": fwd2_ms,
                "fwd_bwd_frac": (6.0 * rows * N * d / ((fwd_ms + bwd_ms) * 1e-3) / 1e12) / peak_tf}

    line = {
        "metric": "InfoNCE fwd+bwd pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K_, "warmup": W_,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("bf16" if in_dtype == torch.bfloat16 else "bf16 (fp32 inputs under autocast)" if use_autocast
                  else "f32 (bf16 hi/lo split, 3 tensor-core terms per product)"),
        "data": "synthetic",
        "config": {"workload": wl["name"], "N": N, "d": d, "temperature": 0.07, "path": args.path,
                   "parallelism": f"rows{world}" if world > 1 else "single",
                   "l2": "256 MiB buffer written between timed iterations (inputs are smaller than L2)"},
        "e2e": e2e, "roofline": roofline, "clocks": clocks,
        # kernels of libcontrastyou_b200.so launched inside the timed region of the `value` leg (cy_launch_count() difference)
        "gpu_launches": int(value_launches),
        "parity": parity,
    }

    # ---- IIC leg (config 3); weak scaling over the batch for world > 1
    if not args.no_iic:
        B, Kc, H, Wd, pad = (IIC_CFG[k] for k in ("B", "K", "H", "W", "pad"))
        g2 = torch.Generator().manual_seed(1 + rank)
        x_host = (2 * torch.randn(B, Kc, H, Wd, generator=g2)).softmax(1).pin_memory()
        y_host = (2 * torch.randn(B, Kc, H, Wd, generator=g2)).softmax(1).pin_memory()
        x_dev, y_dev = x_host.to(dev), y_host.to(dev)
        iic = IIDSegmentationLoss(padding=pad)
        if world > 1:
            cyd.shard_iic_loss(iic)

        def iic_step(xa, ya):
            xa = xa.detach().requires_grad_(); ya = ya.detach().requires_grad_()
            loss = iic(xa, ya)
            loss.backward()
            return loss
        px = B * H * Wd * world
        # the step (~0.14 ms of kernels) is shorter than its own enqueue time in eager PyTorch (~0.16 ms: autograd node, 4 launches),
        # so it is timed twice: at the host's pace, and queued behind a device-side gate (device time; what a training loop with
        # a deep GPU queue sees — the headline, since every rank of a sharded step otherwise waits for the slowest HOST each step)
        iic_host_ms = timed_loop(lambda: iic_step(x_dev, y_dev), W_, K_) / K_
        iic_ms = timed_loop(lambda: iic_step(x_dev, y_dev), W_, K_, gate_ms=min(0.3 * K_, 30.0)) / K_
        iic_launches = counted["launches"]
        # parity of the timed IIC configuration against the C oracle (N=1, full batch), outside the timed region
        iic_parity = None
        if world == 1 and not args.no_cpu:
            from oracle import c_oracle
            c_oracle.set_num_threads(host_cores())
            xa, ya = x_dev.detach().clone().requires_grad_(), y_dev.detach().clone().requires_grad_()
            l_ = iic(xa, ya)
            l_.backward()
            o_ = c_oracle.iic_fwd_bwd(x_host.numpy(), y_host.numpy(), pad, prec=1)
            lr = abs(float(l_.item()) - o_["loss"]) / abs(o_["loss"])
            gr = max(float((xa.grad.cpu() - torch.from_numpy(o_["grad_x"])).abs().max() / abs(o_["grad_x"]).max()),
                     float((ya.grad.cpu() - torch.from_numpy(o_["grad_y"])).abs().max() / abs(o_["grad_y"]).max()))
            iic_parity = {"loss": float(l_.item()), "ref_loss": o_["loss"], "loss_rel": lr, "grad_rel": gr,
                          "against": "C oracle (float64) at config 3's full size", "tolerance": {"loss_rel": 1e-4, "grad_rel": 1e-4},
                          "ok": bool(lr <= 1e-4 and gr <= 1e-4)}
        ne = max(3, K_ // 2)
        iic_e2e_serial_ms = timed_loop(lambda: iic_step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True)).item(),
                                       2, ne) / ne
        pfi = HostPrefetcher([x_host, y_host], dev)
        iic_e2e_ms = timed_loop(lambda: iic_step(*pfi.next()).item(), 2, ne) / ne
        del pfi
        # kernel level
        joint = torch.empty(Kc, Kc, 3, 3, device=dev, dtype=torch.float64)
        wsj_b = lib.cy_iic_workspace_bytes(B, Kc, H, Wd, pad)
        wsj = torch.empty(wsj_b, dtype=torch.uint8, device=dev)
        dj = torch.randn(Kc, Kc, 3, 3, device=dev) * 1e-6
        dxo, dyo = torch.empty_like(x_dev), torch.empty_like(y_dev)
        j_ms = timed_loop(lambda: L.check(lib.cy_iic_joint(x_dev.data_ptr(), y_dev.data_ptr(), 0, B, Kc, H, Wd, pad, joint.data_ptr(),
                                                           wsj.data_ptr(), wsj_b, st), "joint"), 2, kreps) / kreps
        b_ms = timed_loop(lambda: L.check(lib.cy_iic_bwd(x_dev.data_ptr(), y_dev.data_ptr(), 0, B, Kc, H, Wd, pad, dj.data_ptr(),
                                                         one.data_ptr(), dxo.data_ptr(), dyo.data_ptr(), st), "bwd"), 2, kreps) / kreps
        bytes_map = 2 * B * Kc * H * Wd * 4
        gbs = lambda by, ms: by / (ms * 1e-3) / 1e9
        line["iic"] = {
            "metric": "IIC-seg fwd+bwd pixels/s", "value": px / (iic_ms * 1e-3), "unit": "pixels/s", "ms_per_step": iic_ms,
            "host_paced_ms_per_step": iic_host_ms,
            "timing": "K steps queued behind a device-side gate, one CUDA-event pair per step, L2 flushed between steps "
                      "(host_paced_ms_per_step: same without the gate)",
            "scaling": "weak", "dtype": "f32",
            "config": {"workload": f"cfg3 IIDSegmentationLoss K={Kc} padding={pad} batch {B}x{H}x{Wd} per GPU"},
            "e2e": {"value": px / (iic_e2e_ms * 1e-3), "unit": "pixels/s", "h2d_bytes_per_step": bytes_map, "d2h_bytes_per_step": 4,
                    "ms_per_step": iic_e2e_ms, "serial_ms_per_step": iic_e2e_serial_ms,
                    "pipeline": "double-buffered H2D on a copy stream (HostPrefetcher); PCIe-bound: 128 MB per step"},
            "roofline": {"bound": "hbm", "kernel": "cy_iic_joint + cy_iic_bwd (3 x 2*B*K*H*W*4 B algorithmic)",
                         "achieved": gbs(3 * bytes_map, j_ms + b_ms), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs(3 * bytes_map, j_ms + b_ms) / peaks["hbm_gbs"],
                         "traffic": (traffic.get("iic_joint_mma_kernel<2,4,1,2> cfg3", {}).get("bytes", 0)
                                     + traffic.get("iic_bwd_tc_kernel<5> cfg3", {}).get("bytes", 0)) or None,
                         "joint": {"ms": j_ms, "achieved": gbs(bytes_map, j_ms), "frac": gbs(bytes_map, j_ms) / peaks["hbm_gbs"]},
                         "bwd": {"ms": b_ms, "achieved": gbs(2 * bytes_map, b_ms), "frac": gbs(2 * bytes_map, b_ms) / peaks["hbm_gbs"]},
                         "peak_source": peaks["source"] + " copy bandwidth"},
            "gpu_launches": int(iic_launches), "parity": iic_parity,
        }

        # ---- sub-head stack: the hooks' python sum over S sub-head pairs vs forward_heads (one joint, one
        # reduction, one epilogue and one adjoint launch for all heads), at a cluster-head shape and at the cfg3 map size
        if world == 1:
            heads = {}
            for tag, (S_, Bh, Hh, Wh) in (("5 heads x 8x10x64x64", (5, 8, 64, 64)), ("3 heads x 16x10x224x224", (3, 16, 224, 224))):
                gh = torch.Generator(device=dev).manual_seed(5)
                hx = [(2 * torch.randn(Bh, Kc, Hh, Wh, device=dev, generator=gh)).softmax(1) for _ in range(S_)]
                hy = [(2 * torch.randn(Bh, Kc, Hh, Wh, device=dev, generator=gh)).softmax(1) for _ in range(S_)]
                crit_h = IIDSegmentationLoss(padding=pad)

                def loop_step():
                    xs = [t.detach().requires_grad_() for t in hx]; ys = [t.detach().requires_grad_() for t in hy]
                    l_ = sum(crit_h(a, b) for a, b in zip(xs, ys)) / S_
                    l_.backward()
                    return l_

                def heads_step():
                    xs = [t.detach().requires_grad_() for t in hx]; ys = [t.detach().requires_grad_() for t in hy]
                    l_ = crit_h.forward_heads(xs, ys)
                    l_.backward()
                    return l_
                loop_ms = timed_loop(loop_step, W_, K_) / K_
                loop_launches = counted["launches"] // K_
                heads_ms = timed_loop(heads_step, W_, K_) / K_
                heads[tag] = {"python_loop_ms": loop_ms, "forward_heads_ms": heads_ms, "launches_per_step": [loop_launches,
                              counted["launches"] // K_], "loss_rel": abs(float(heads_step()) - float(loop_step())) / abs(float(loop_step()))}
                del hx, hy
            line["iic"]["sub_heads"] = heads

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"], _ = cpu_supcon_baseline(steps=3, warmup=1)
        if "iic" in line:
            line["iic"]["cpu_baseline"] = cpu_iic_baseline(steps=3, warmup=1)
    if graphed is not None:
        line["graphed"] = graphed
    if rs_variant is not None:
        line["reduce_scatter_design"] = rs_variant
    if world > 1:
        line["config"]["exchange"] = ("peer memory (cy_p2p_push + symmetric-memory signal barriers)"
                                      if getattr(crit, "_px", None) is not None else "NCCL all-gathers")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        if args.graph:      # tearing the NCCL communicator down while captured graphs still reference it blocks: leave it to exit
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
