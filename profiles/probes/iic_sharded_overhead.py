"""Where does the batch-sharded IIC step spend its time?  (torchrun, >= 2 ranks.)

For the sharded and the unsharded criterion on the bench's IIC workload: (A) host enqueue time per step (perf_counter over a
loop without synchronising), (B) device time per step over a back-to-back loop (one event pair around 50 steps), (C) the
bench's method (L2 flush + one event pair per step).  B < A means the step is host-bound."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import contrast_you_b200.distributed as cyd  # noqa: E402
from contrast_you_b200.losses.discreteMI import IIDSegmentationLoss  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B, K, H, W, pad = 32, 10, 224, 224, 1
    g = torch.Generator().manual_seed(1 + rank)
    x = (2 * torch.randn(B, K, H, W, generator=g)).softmax(1).to(dev)
    y = (2 * torch.randn(B, K, H, W, generator=g)).softmax(1).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    plain = IIDSegmentationLoss(padding=pad)
    sharded = IIDSegmentationLoss(padding=pad)
    cyd.shard_iic_loss(sharded)
    nccl = IIDSegmentationLoss(padding=pad)
    cyd.shard_iic_loss(nccl, exchange="nccl")

    def step(crit):
        xa, ya = x.detach().requires_grad_(), y.detach().requires_grad_()
        loss = crit(xa, ya)
        loss.backward()
        return loss

    n = 50
    for name, crit in (("unsharded", plain), ("sharded p2p", sharded), ("sharded nccl", nccl)):
        for _ in range(5):
            step(crit)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            step(crit)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        host_us, dev_us = (t1 - t0) / n * 1e6, e0.elapsed_time(e1) / n * 1e3
        dist.barrier(); torch.cuda.synchronize()
        evs = []
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(crit); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        per = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        print(f"[rank {rank}] {name:13s}: host enqueue {host_us:7.1f} us/step | back-to-back device {dev_us:7.1f} us/step | "
              f"flushed per-step events: median {per[n // 2]:7.1f}  p10 {per[n // 10]:7.1f}  p90 {per[9 * n // 10]:7.1f}  "
              f"mean {sum(per) / n:7.1f} us", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
