"""CUDA-graph replay of the row-sharded InfoNCE step (collectives included), launched under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/multi_gpu_graph_check.py [--n-loc 4096] [--steps 20]

``ShardedSupConLoss(deferred_checks=True)`` has no host read in its forward, so forward and backward (label all-gather, local
sort, pack, Z all-gather, forward strip, statistics all-gather, backward strip, unpack) can be captured once with
``torch.cuda.make_graphed_callables`` and replayed.  The script checks replay == eager on fresh inputs and reports the
device time per step of both, max over ranks.  Kept apart from multi_gpu_check.py: NCCL under stream capture is the one
part of the stack that depends on the NCCL / driver pairing of the box.

Status (round 1, B200 pool, torch 2.11 / NCCL 2.28.9): one rank: replay bit-identical to eager, 0.586 -> 0.449 ms/step at
N = 16384; two ranks at N = 8192: same loss, bf16 gradients within an ulp, 0.732 -> 0.219 ms/step.  Caveat found on the way:
`dist.destroy_process_group()` blocks while captured graphs reference the communicator, so the script (and `bench.py
--graph`) leave the teardown to process exit.  Run under `timeout`."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrast_you_b200 import distributed as cyd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-loc", type=int, default=4096, help="samples per view and rank")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--watchdog", type=int, default=0, help="dump all python stacks and exit after this many seconds")
    args = ap.parse_args()
    if args.watchdog:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_loc, d = args.n_loc, 256
    g = torch.Generator().manual_seed(100 + rank)

    def batch():
        f1 = torch.nn.functional.normalize(torch.randn(n_loc, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
        f2 = torch.nn.functional.normalize(torch.randn(n_loc, d, generator=g), dim=1).to(torch.bfloat16).to(dev)
        lab = torch.randint(0, max(4, n_loc * world // 16), (n_loc,), generator=g).to(torch.int32).to(dev)   # int32: no torch.unique
        return f1, f2, lab

    crit = cyd.ShardedSupConLoss(deferred_checks=True)

    def fn(a, b, lab):
        return crit(a, b, target=lab)

    def eager(f1, f2, lab):
        a, b = f1.detach().requires_grad_(), f2.detach().requires_grad_()
        loss = fn(a, b, lab)
        loss.backward()
        return loss.detach().clone(), a.grad.clone(), b.grad.clone()

    f1, f2, lab = batch()
    eager(f1, f2, lab)                                  # communicator set-up and workspace allocation outside the capture
    torch.cuda.synchronize()
    dist.barrier()
    sa, sb, sl = f1.clone().requires_grad_(), f2.clone().requires_grad_(), lab.clone()
    graphed = torch.cuda.make_graphed_callables(fn, (sa, sb, sl))

    def replay(f1, f2, lab):
        sa.detach().copy_(f1)
        sb.detach().copy_(f2)
        sl.copy_(lab)
        sa.grad = None
        sb.grad = None
        loss = graphed(sa, sb, sl)
        loss.backward()
        return loss.detach().clone(), sa.grad.clone(), sb.grad.clone()

    ok = True
    for it in range(3):
        f1, f2, lab = batch()
        le, gae, gbe = eager(f1, f2, lab)
        lg, gag, gbg = replay(f1, f2, lab)
        e_loss = abs(le.item() - lg.item()) / abs(le.item())
        e_grad = max(((gae.float() - gag.float()).abs().max() / gae.float().abs().max()).item(),
                     ((gbe.float() - gbg.float()).abs().max() / gbe.float().abs().max()).item())
        good = e_loss < 1e-6 and e_grad < 2e-2          # same kernels, same inputs: the bf16 gradients differ by at most an ulp
                                                        # (fp32 atomics of the split backward commit in a different order)
        ok &= good
        print(f"[rank {rank}] batch {it}: eager {le.item():.6f} graph {lg.item():.6f} rel {e_loss:.1e} grad rel {e_grad:.1e} "
              f"{'OK' if good else 'FAIL'}", flush=True)
    crit.raise_if_flagged()

    def timed(step):
        f1, f2, lab = batch()
        for _ in range(3):
            step(f1, f2, lab)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(f1, f2, lab)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_eager, t_graph = timed(eager), timed(replay)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        N = 2 * n_loc * world
        print(f"N = {N} on {world} GPUs: eager {t_eager:.3f} ms/step ({N * N / t_eager / 1e9:.1f}e12 pairs/s), "
              f"graph replay {t_graph:.3f} ms/step ({N * N / t_graph / 1e9:.1f}e12 pairs/s)")
        print("graph replay parity OK" if flag.item() == 1.0 else "graph replay parity FAILED")
    # destroy_process_group() blocks while captured graphs still reference the communicator: leave the teardown to exit
    code = 0 if flag.item() == 1.0 else 1
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(code)


if __name__ == "__main__":
    main()
