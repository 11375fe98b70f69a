"""Multi-GPU parity check, launched under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Row-sharded global-batch InfoNCE (ShardedSupConLoss) and batch-sharded IIC (shard_iic_loss) on G ranks must reproduce
the single-process modules on the concatenated batch: same loss on every rank, and each rank's input gradients equal
the corresponding rows / images of the single-process gradients."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_you_b200.losses import SupConLoss1, IIDSegmentationLoss  # noqa: E402
from contrast_you_b200 import distributed as cyd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = [(64, torch.float32, "simt", 1e-4, {}), (512, torch.bfloat16, "tcgen05", 2e-2, {}), (2048, torch.bfloat16, "auto", 2e-2, {}),
             # a local batch whose row block is not 128-aligned (2 x 96 = 192 rows) must fall back, not raise
             (96, torch.bfloat16, "auto", 2e-2, {}),
             # the rest of the family, sharded; and the north_star's reduce-scatter design (i) next to the default (ii)
             (1024, torch.bfloat16, "auto", 2e-2, {"exclude_other_pos": True}),
             (512, torch.bfloat16, "auto", 2e-2, {"backward_design": "reduce_scatter"})]
    for (n_loc, dtype, path, tol, extra) in cases:
        g = torch.Generator().manual_seed(1234)
        n = n_loc * world
        f1 = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dtype)
        f2 = torch.nn.functional.normalize(torch.randn(n, 256, generator=g), dim=1).to(dtype)
        lab = torch.randint(0, max(4, n // 16), (n,), generator=g)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        a = f1[sl].to(dev).requires_grad_()
        b = f2[sl].to(dev).requires_grad_()
        crit = cyd.ShardedSupConLoss(path=path, **extra)
        loss = crit(a, b, target=lab[sl].tolist())
        loss.backward()
        # second call with DEVICE int64 labels (the cached-tensor route): must give the same numbers
        a2, b2 = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
        lab_dev = lab[sl].to(dev)
        for _ in range(2):
            a2.grad = None
            loss2 = crit(a2, b2, target=lab_dev)
            loss2.backward()
        # single-process reference on the whole batch (every rank recomputes it; SIMT fp32 path for the bf16 cases too)
        A = f1.to(dev).requires_grad_()
        B = f2.to(dev).requires_grad_()
        ref = SupConLoss1(path="simt", exclude_other_pos=bool(extra.get("exclude_other_pos")))(A, B, target=lab.tolist())
        ref.backward()
        e_loss = abs(loss.item() - ref.item()) / abs(ref.item())
        ga, gA = a.grad.float(), A.grad[sl].float()
        e_grad = ((ga - gA).abs().max() / gA.abs().max()).item()
        same = abs(loss2.item() - loss.item()) <= 1e-6 * abs(loss.item()) and torch.equal(a2.grad, a.grad)
        good = e_loss < tol and e_grad < tol and same
        ok &= good
        print(f"[rank {rank}] infonce n_loc={n_loc} {dtype} {path} {extra}: loss {loss.item():.6f} ref {ref.item():.6f} "
              f"rel {e_loss:.2e} grad rel {e_grad:.2e} repeat-identical {same} {'OK' if good else 'FAIL'}", flush=True)

    g = torch.Generator().manual_seed(99)
    Bt, K, H, W = 2 * world, 10, 48, 40
    x = torch.randn(Bt, K, H, W, generator=g).softmax(1)
    y = torch.randn(Bt, K, H, W, generator=g).softmax(1)
    sl = slice(rank * 2, rank * 2 + 2)
    for pad, sym in ((1, False), (0, True)):
        xs, ys = x[sl].to(dev).requires_grad_(), y[sl].to(dev).requires_grad_()
        crit = cyd.shard_iic_loss(IIDSegmentationLoss(padding=pad, symmetric=sym))
        loss = crit(xs, ys)
        loss.backward()
        X, Y = x.to(dev).requires_grad_(), y.to(dev).requires_grad_()
        ref = IIDSegmentationLoss(padding=pad, symmetric=sym)(X, Y)
        ref.backward()
        e_loss = abs(loss.item() - ref.item()) / abs(ref.item())
        e_grad = ((xs.grad - X.grad[sl]).abs().max() / X.grad.abs().max()).item()
        good = e_loss < 1e-4 and e_grad < 1e-4      # the fp32 parity bar (partial joints are summed in a different order)
        ok &= good
        print(f"[rank {rank}] iic pad={pad} sym={sym}: loss {loss.item():.7f} ref {ref.item():.7f} rel {e_loss:.2e} "
              f"grad rel {e_grad:.2e} {'OK' if good else 'FAIL'}", flush=True)
    # the sub-head stack under sharding: forward_heads exchanges ALL heads' partial joints as one array ([world][S][K,K,T,T]
    # peer slots, or one NCCL all-reduce) and must equal the mean of single-process criteria on the gathered batches; from
    # probabilities and from logits (fused SoftmaxWithT)
    S = 3
    lx = [2 * torch.randn(Bt, K, H, W, generator=g) for _ in range(S)]
    ly = [2 * torch.randn(Bt, K, H, W, generator=g) for _ in range(S)]
    for exchange in ("auto", "nccl"):
        for from_logits in (False, True):
            crit = cyd.shard_iic_loss(IIDSegmentationLoss(padding=1), exchange=exchange)
            loc_x = [t[sl].to(dev).requires_grad_() for t in lx]
            loc_y = [t[sl].to(dev).requires_grad_() for t in ly]
            if from_logits:
                loss = crit.forward_heads(loc_x, loc_y, logits_T=0.5)
            else:
                loss = crit.forward_heads([t.softmax(1) for t in loc_x], [t.softmax(1) for t in loc_y])
            loss.backward()
            X = [t.to(dev).requires_grad_() for t in lx]
            Y = [t.to(dev).requires_grad_() for t in ly]
            Tm = 0.5 if from_logits else 1.0
            plain = IIDSegmentationLoss(padding=1)
            ref = sum(plain(torch.softmax(a / Tm, 1), torch.softmax(b / Tm, 1)) for a, b in zip(X, Y)) / S
            ref.backward()
            e_loss = abs(loss.item() - ref.item()) / abs(ref.item())
            e_grad = max(((a.grad - A.grad[sl]).abs().max() / A.grad.abs().max()).item() for a, A in zip(loc_x + loc_y, X + Y))
            good = e_loss < 1e-4 and e_grad < 1e-4
            ok &= good
            print(f"[rank {rank}] iic forward_heads S={S} exchange={exchange} logits={from_logits}: loss {loss.item():.7f} ref "
                  f"{ref.item():.7f} rel {e_loss:.2e} grad rel {e_grad:.2e} {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if flag.item() != 1.0:
        sys.exit(1)
    if rank == 0:
        print("multi-GPU parity OK")


if __name__ == "__main__":
    main()
