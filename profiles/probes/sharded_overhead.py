"""Per-step constant of the sharded InfoNCE step: GPU time (CUDA events) and wall time of the default (one host read per step),
the deferred-checks (no host read) and the graph-replayed step, next to the sum of the two strip kernels.
torchrun --nproc-per-node 2 profiles/probes/sharded_overhead.py [N]   (N = 32768 at 2 ranks gives the per-rank strip of config 4 at 8)"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from contrast_you_b200 import distributed as cyd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = 256
n_loc = N // 2 // world
g = torch.Generator().manual_seed(rank)
f1 = torch.nn.functional.normalize(torch.randn(n_loc, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
f2 = torch.nn.functional.normalize(torch.randn(n_loc, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
lab = torch.randint(0, 4096, (n_loc,), generator=g).to(torch.int32).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def measure(name, crit, steps=30, do_flush=True):
    def step():
        a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_()
        loss = crit(a, b, target=lab); loss.backward(); return loss
    for _ in range(5): step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    evs = []
    t0 = time.perf_counter()
    for _ in range(steps):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps * 1e3
    gpu = sum(a.elapsed_time(b) for a, b in evs) / steps
    # CPU time to ISSUE a step when nothing blocks: deferred mode only
    t = torch.tensor([gpu, wall], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{name:34s} world {world} N {N}: GPU {t[0].item():.3f} ms/step (events, max over ranks)  wall {t[1].item():.3f} ms/step", flush=True)


measure("default (host read per step)", cyd.ShardedSupConLoss())
measure("default, no L2 flush", cyd.ShardedSupConLoss(), do_flush=False)
measure("deferred_checks", cyd.ShardedSupConLoss(deferred_checks=True))
measure("deferred_checks, nccl exchange", cyd.ShardedSupConLoss(deferred_checks=True, exchange="nccl"))
measure("default, nccl exchange", cyd.ShardedSupConLoss(exchange="nccl"))
# CPU issue time of one step (deferred: nothing blocks the host)
crit = cyd.ShardedSupConLoss(deferred_checks=True)
for _ in range(3):
    a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_(); crit(a, b, target=lab).backward()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_(); crit(a, b, target=lab).backward()
cpu = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
if rank == 0: print(f"CPU issue time of a deferred step: {cpu:.3f} ms (the GPU queue absorbs it while it is below the GPU time)")
dist.destroy_process_group()
