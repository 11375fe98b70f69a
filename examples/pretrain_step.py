#!/usr/bin/env python
"""BASELINE.json config 5 — end-to-end encoder-pretrain step with the InfoNCE and discrete-MI hooks, on synthetic
ACDC-shaped batches, single GPU or DDP (torchrun).

This is a STAND-IN for the reference's Trainer / Epocher / hook stack (out of scope for this repo, SURVEY.md §2, and
not importable here: §8c): it follows the protocol of ``semi_seg/epochers/pretrain.py:93-98`` and the two hooks
(``semi_seg/hooks/infonce.py:222-245`` — global InfoNCE at Conv5 with partition labels; ``discretemi.py:99-113`` —
IIC at Up_conv2 averaged over the sub-heads) so that the drop-in loss modules are exercised the way the hooks call them:
one forward to Up_conv2 with feature taps, projector heads, criteria, ``.item()`` metering, AMP, optimizer step.
The network is a plain torch U-Net with the reference's layer widths (``arch/unet.py:50-110``: max_channel/16 x
{1,2,4,8,16}); its convolutions stay on stock torch/cuDNN by contract.

    python examples/pretrain_step.py --steps 10                       # one GPU
    torchrun --nproc-per-node 8 examples/pretrain_step.py --steps 10  # DDP, 18 scans x 2 views per rank
Prints one JSON line (rank 0): images/s over all ranks, ms/step, and the share of the step spent in the two criteria.

Under DDP the two criteria are the GLOBAL-BATCH forms (``--local-criteria`` restores what the reference would do under a
DDP launch: every rank contrasts its own 18 images): ``ShardedSupConLoss(grad_scale=world)`` and a batch-sharded
``IIDSegmentationLoss`` whose loss is scaled by ``world`` the same way, so that DDP's gradient AVERAGE over ranks equals the
single-process gradient of the global-batch loss.  ``--check`` verifies exactly that on the GPU box: after one backward the
DDP-averaged parameter gradients are compared with a single-process evaluation of the concatenated batch of all ranks on
the same weights (BatchNorm in eval mode for the check: batch statistics are per-rank under DDP by design).
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from contrast_you_b200.labels import PartitionLabelGenerator  # noqa: E402
from contrast_you_b200.losses import IIDSegmentationLoss, SupConLoss1  # noqa: E402
from contrast_you_b200.projectors import DenseClusterHead, ProjectionHead  # noqa: E402


def _block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                         nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class StandInUNet(nn.Module):
    """5-level U-Net; ``forward`` returns the taps the two hooks read: Conv5 (bottleneck) and Up_conv2 (full resolution)."""

    def __init__(self, input_dim=1, max_channel=512):
        super().__init__()
        w = [max_channel // 16 * m for m in (1, 2, 4, 8, 16)]
        self.enc = nn.ModuleList([_block(input_dim if i == 0 else w[i - 1], w[i]) for i in range(5)])
        self.up = nn.ModuleList([nn.Sequential(nn.Upsample(scale_factor=2), nn.Conv2d(w[i], w[i - 1], 3, padding=1, bias=False),
                                               nn.BatchNorm2d(w[i - 1]), nn.ReLU(inplace=True)) for i in range(4, 0, -1)])
        self.dec = nn.ModuleList([_block(2 * w[i - 1], w[i - 1]) for i in range(4, 0, -1)])
        self.widths = w

    def forward(self, x):
        skips = []
        for i, blk in enumerate(self.enc):
            x = blk(x if i == 0 else nn.functional.max_pool2d(x, 2))
            skips.append(x)
        conv5 = x
        for up, dec, skip in zip(self.up, self.dec, reversed(skips[:-1])):
            x = dec(torch.cat((skip, up(x)), dim=1))
        return conv5, x


class PretrainModel(nn.Module):
    """network + the two projector heads in one module, so that stock DistributedDataParallel wraps every parameter the
    hooks train (in the reference the heads are parameters of TrainerHook modules inside the Trainer's module tree)."""

    def __init__(self, max_channel, clusters, subheads):
        super().__init__()
        self.net = StandInUNet(1, max_channel)
        w = self.net.widths
        self.infonce_head = ProjectionHead(input_dim=w[4], hidden_dim=256, output_dim=256, head_type="mlp", normalize=True)
        self.mi_head = DenseClusterHead(input_dim=w[0], num_clusters=clusters, num_subheads=subheads, head_type="linear")

    def forward(self, x):
        conv5, up2 = self.net(x)
        # fused_softmax: the cluster head stops before SoftmaxWithT and the criterion takes the logits (forward_heads(..., logits_T))
        return self.infonce_head(conv5), self.mi_head(up2, skip_softmax=getattr(self, "fused_softmax", False))


def ddp_gradient_check(model, fwd, infonce, mi, mi_scale, img, img_tf, labels, dev, rank, world):
    """one fp32 backward through DDP with the global-batch criteria vs a single-process backward on the concatenated batch
    of all ranks, same weights, BatchNorm in eval mode.  Returns the max relative parameter-gradient error (over ranks)."""
    model.eval()
    fused, model.fused_softmax = getattr(model, "fused_softmax", False), False      # the check compares the plain modules
    x = torch.cat((img.to(dev), img_tf.to(dev)))
    model.zero_grad(set_to_none=True)
    z, probs = fwd(x)
    z1, z2 = torch.chunk(z, 2, dim=0)
    pairs = [torch.chunk(p.float(), 2, 0) for p in probs]
    loss = infonce(z1, z2, target=labels) + (0.1 * mi_scale) * sum(mi(a, b) for a, b in pairs) / len(pairs)
    loss.backward()                                                       # DDP averages the parameter gradients over ranks
    got = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    # single process: gather every rank's images, evaluate the plain modules on the whole batch with the same weights
    def gather(t):
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t.to(dev).contiguous())
        return out.flatten(0, 1)
    xa, xb = gather(img), gather(img_tf)
    all_labels = labels * world                                            # every rank holds the same partition list
    model.zero_grad(set_to_none=True)
    z, probs = model(torch.cat((xa, xb)))
    z1, z2 = torch.chunk(z, 2, dim=0)
    pairs = [torch.chunk(p.float(), 2, 0) for p in probs]
    ref = SupConLoss1()(z1, z2, target=all_labels) + 0.1 * sum(IIDSegmentationLoss(padding=1)(a, b) for a, b in pairs) / len(pairs)
    ref.backward()
    # per-parameter error relative to that parameter's largest gradient entry, and the error of the whole gradient vector
    # relative to its norm (cuDNN picks different algorithms for a 36- and a 72-image batch: parameters whose gradient is a
    # near-total cancellation carry fp32 summation noise of their own size, the global figure does not)
    worst, worst_name, num, den = 0.0, "", 0.0, 0.0
    for n, p in model.named_parameters():
        if p.grad is None or n not in got:
            continue
        diff = (got[n] - p.grad).double()
        num += float((diff * diff).sum())
        den += float((p.grad.double() ** 2).sum())
        scale = float(p.grad.abs().max())
        if scale > 0:
            e = float(diff.abs().max()) / scale
            if e > worst:
                worst, worst_name = e, f"{n} (|grad|max {scale:.2e})"
    glob = (num / den) ** 0.5 if den > 0 else float("nan")
    t = torch.tensor([worst, glob, abs(float(loss.item()) / world - float(ref.item()))], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    model.zero_grad(set_to_none=True)
    model.train()
    model.fused_softmax = fused
    return {"global_grad_rel_err": float(t[1]), "max_param_grad_rel_err": float(t[0]), "worst_param": worst_name,
            "loss_abs_diff": float(t[2]), "ok": bool(t[1] < 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scans", type=int, default=6, help="scans per rank; x3 partitions = images per view (config/pretrain.yaml:15)")
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--max-channel", type=int, default=512)
    ap.add_argument("--clusters", type=int, default=10)
    ap.add_argument("--subheads", type=int, default=5)
    ap.add_argument("--no-amp", action="store_true")
    ap.add_argument("--batched-heads", action="store_true", help="IIDSegmentationLoss.forward_heads instead of the python sum")
    ap.add_argument("--fused-softmax", action="store_true",
                    help="cluster head returns logits; SoftmaxWithT runs inside the criterion (one streaming launch forward, "
                         "backward in the IIC adjoint's epilogue); implies --batched-heads")
    ap.add_argument("--local-criteria", action="store_true", help="under DDP: per-rank criteria (no exchange), as the reference would run")
    ap.add_argument("--check", action="store_true", help="DDP global-batch gradients == single-process gradients on the concatenated batch")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)

    model = PretrainModel(args.max_channel, args.clusters, args.subheads).to(dev)
    model.fused_softmax = bool(args.fused_softmax)
    mi_temperature = float(model.mi_head.temperature)
    fwd = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if world > 1 else model
    global_batch = world > 1 and not args.local_criteria
    if global_batch:
        from contrast_you_b200 import distributed as cyd
        infonce = cyd.ShardedSupConLoss(grad_scale=world)
        mi = cyd.shard_iic_loss(IIDSegmentationLoss(padding=1))
    else:
        infonce, mi = SupConLoss1(), IIDSegmentationLoss(padding=1)          # the drop-in criteria (no parameters)
    mi_scale = float(world) if global_batch else 1.0
    opt = torch.optim.Adam(model.parameters(), lr=1e-6)
    amp = not args.no_amp
    scaler = torch.amp.GradScaler("cuda", enabled=amp)

    n_img = args.scans * 3
    gen = torch.Generator().manual_seed(rank)
    img = torch.rand(n_img, 1, args.size, args.size, generator=gen).pin_memory()
    img_tf = torch.rand(n_img, 1, args.size, args.size, generator=gen).pin_memory()
    partitions = [f"{p}" for _ in range(args.scans) for p in ("0", "1", "2")]
    labels = PartitionLabelGenerator()(partitions)                       # semi_seg/epochers/helper.py:54-58
    t_crit = [0.0]

    def step():
        x = torch.cat((img.to(dev, non_blocking=True), img_tf.to(dev, non_blocking=True)))
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            z, probs = fwd(x)
            c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
            c0.record()
            # --- InfoNCE hook (infonce.py:222-245): chunk the two views, partition labels
            z1, z2 = torch.chunk(z, 2, dim=0)
            loss_nce = infonce(z1, z2, target=labels)
            # --- discrete-MI hook (discretemi.py:99-113): sub-head list, chunk, mean of the criterion over the heads
            pairs = [torch.chunk(p.float(), 2, 0) for p in probs]
            if args.fused_softmax:
                loss_mi = mi.forward_heads([a for a, _ in pairs], [b for _, b in pairs], logits_T=mi_temperature)
            elif args.batched_heads:
                loss_mi = mi.forward_heads([a for a, _ in pairs], [b for _, b in pairs])
            else:
                loss_mi = sum(mi(a, b) for a, b in pairs) / len(pairs)
            c1.record()
            loss = loss_nce + (0.1 * mi_scale) * loss_mi
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        meters = (loss_nce.item(), loss_mi.item())                       # hooks meter with .item() every batch
        t_crit[0] += c0.elapsed_time(c1)
        return meters

    check = None
    if args.check and world > 1:
        check = ddp_gradient_check(model, fwd, infonce, mi, mi_scale, img, img_tf, labels, dev, rank, world)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_crit[0] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        meters = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "cfg5 stand-in pretrain step, images/s", "value": 2 * n_img * world / (ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "ms_per_step": ms, "criteria_fwd_ms_per_step": t_crit[0] / args.steps,
            "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / args.steps, "amp": amp,
            "config": {"images_per_rank": 2 * n_img, "size": args.size, "max_channel": args.max_channel,
                       "clusters": args.clusters, "subheads": args.subheads, "padding": 1,
                       "iic_heads": "fused softmax (logits)" if args.fused_softmax else ("batched" if args.batched_heads else "python loop")},
            "criteria": "global-batch (row-sharded InfoNCE + batch-sharded IIC)" if global_batch else "per-rank",
            "ddp_gradient_check": check,
            "last_losses": {"infonce": meters[0], "discrete_mi": meters[1]}, "data": "synthetic"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
