"""World-size-2 `gloo` tests (CPU) of the multi-GPU plumbing in contrast_you_b200/distributed.py.

The CUDA kernels cannot run here, so each rank evaluates ITS shard with the CPU oracle (test infrastructure) and the
product's exchange code (gather_rank_major / rank_major_labels / make_stats_exchange / make_joint_reduce) has to turn
the per-rank pieces into exactly the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _canon_cpu(raw, n):
    """CPU stand-in for cy_labels_canonicalize(kind=float32): bit pattern of (x + 0.0), tiled over the two views"""
    bits = (raw.to(torch.float32) + 0.0).view(torch.int32)
    return torch.cat([bits, bits])


def _worker(rank, port, results):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from contrast_you_b200 import distributed as cyd
        from contrast_you_b200 import _lib as L
        from oracle import contrastive_np as OC, discrete_mi_np as OM

        # ---------------- row-sharded InfoNCE
        g = torch.Generator().manual_seed(7)
        n_loc, d, t = 12, 16, 0.07
        n = n_loc * WORLD
        f1 = torch.nn.functional.normalize(torch.randn(n, d, generator=g, dtype=torch.float64), dim=1)
        f2 = torch.nn.functional.normalize(torch.randn(n, d, generator=g, dtype=torch.float64), dim=1)
        lab = torch.randint(0, 4, (n,), generator=g)
        sl = slice(rank * n_loc, (rank + 1) * n_loc)
        local = torch.cat([f1[sl], f2[sl]]).requires_grad_()
        z_all = cyd.gather_rank_major(local)                                   # [G * 2 n_loc, d], rank-major
        raw_all = torch.empty(n, dtype=torch.float32)
        dist.all_gather_into_tensor(raw_all, lab[sl].to(torch.float32))
        labels = cyd.rank_major_labels(raw_all, WORLD, _canon_cpu)
        rb, re = cyd.row_range(n_loc)
        assert (rb, re) == (rank * 2 * n_loc, (rank + 1) * 2 * n_loc)
        N = 2 * n
        # this rank's strip with the oracle's closed form: statistics of the owned rows only
        Z = z_all.detach().numpy()
        lb = labels.numpy()
        S = Z[rb:re] @ Z.T / t
        m = 1.0 / t
        rows = np.arange(rb, re)
        P = (lb[rb:re, None] == lb[None, :]).astype(np.float64); P[rows - rb, rows] = 0
        E = np.exp(S - m); E[rows - rb, rows] = 0
        D, c = E.sum(1), P.sum(1)
        stats = torch.zeros(L.CY_NSTAT, N, dtype=torch.float64)
        stats[L.CY_STAT_LOGDEN, rb:re] = torch.from_numpy(np.log(D + 1e-16))
        stats[L.CY_STAT_INVC, rb:re] = torch.from_numpy(1 / c)
        stats[L.CY_STAT_COEF, rb:re] = torch.from_numpy(1 / (D + 1e-16))
        out4 = torch.zeros(4, dtype=torch.float64)
        out4[0] = float(-((P * (S - m)).sum(1) / c - np.log(D + 1e-16)).sum() / N)
        stats_b, out4_b = stats.clone(), out4.clone()
        cyd.make_stats_exchange(n_loc)(stats, out4)                            # the product's exchange step (5 collectives)
        out4_b = cyd.exchange_strip_stats(stats_b, out4_b, rb, re)             # ... and its one-collective form
        assert torch.equal(stats_b, stats) and torch.allclose(out4_b, out4, rtol=1e-15, atol=0)
        # local label sort used by ShardedSupConLoss: sorting every rank's block leaves the loss unchanged and the owned
        # block's permutation maps sorted rows back to the original ones
        sorted_blocks, perm = labels.view(WORLD, 2 * n_loc).sort(dim=1)
        assert torch.equal(labels.view(WORLD, 2 * n_loc)[rank][perm[rank]], sorted_blocks[rank])
        coef, invc = stats[L.CY_STAT_COEF].numpy(), stats[L.CY_STAT_INVC].numpy()
        W = E * (coef[rb:re, None] + coef[None, :]) - P * (invc[rb:re, None] + invc[None, :])
        dz_all = torch.zeros(N, d, dtype=torch.float64)
        dz_all[rb:re] = torch.from_numpy(W @ Z / (t * N))
        z_all.backward(dz_all)                                                 # through _GatherRows.backward
        # single-process truth: the literal oracle on the concatenated batch
        ref = OC.supcon(f1.numpy(), f2.numpy(), target=lab.tolist(), t=t)
        assert abs(out4[0].item() - ref["loss"]) < 1e-12 * abs(ref["loss"]), (out4[0].item(), ref["loss"])
        got = local.grad.numpy()
        want = np.concatenate([ref["grad_f1"][sl], ref["grad_f2"][sl]])
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()

        # ---------------- batch-sharded IIC: all-reduce of the raw joint before the epilogue
        g = torch.Generator().manual_seed(8)
        x = torch.randn(2 * WORLD, 5, 9, 11, generator=g, dtype=torch.float64).softmax(1).numpy()
        y = torch.randn(2 * WORLD, 5, 9, 11, generator=g, dtype=torch.float64).softmax(1).numpy()
        bs = slice(2 * rank, 2 * rank + 2)
        J = torch.from_numpy(OM.raw_joint_2d(x[bs], y[bs], 1))
        npx = cyd.make_joint_reduce()(J, float(2 * 9 * 11))
        assert npx == 2 * WORLD * 9 * 11
        np.testing.assert_allclose(J.numpy(), OM.raw_joint_2d(x, y, 1), rtol=1e-13)
        results[rank] = "ok"
    except Exception as e:  # noqa
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_sharding_plumbing_gloo_world2():
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(port, results), nprocs=WORLD, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}, "\n".join(f"[rank {k}] {v}" for k, v in results.items())
