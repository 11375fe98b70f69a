"""cProfile of the host side of the (sharded) IIC step on rank 0 (torchrun, 2 ranks): which python / torch calls make up the
per-step enqueue time."""
import cProfile
import os
import pstats
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import contrast_you_b200.distributed as cyd  # noqa: E402
from contrast_you_b200.losses.discreteMI import IIDSegmentationLoss  # noqa: E402


def main():
    rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B, K, H, W, pad = 32, 10, 224, 224, 1
    g = torch.Generator().manual_seed(1 + rank)
    x = (2 * torch.randn(B, K, H, W, generator=g)).softmax(1).to(dev)
    y = (2 * torch.randn(B, K, H, W, generator=g)).softmax(1).to(dev)
    crit = IIDSegmentationLoss(padding=pad)
    if os.environ.get("SHARD", "1") == "1":
        cyd.shard_iic_loss(crit)

    def step():
        xa, ya = x.detach().requires_grad_(), y.detach().requires_grad_()
        loss = crit(xa, ya)
        loss.backward()

    for _ in range(10):
        step()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    n = 300
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        step()
    pr.disable()
    torch.cuda.synchronize()
    if rank == 0:
        st = pstats.Stats(pr, stream=sys.stdout)
        st.sort_stats("tottime").print_stats(28)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
