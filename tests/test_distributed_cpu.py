"""World-size-2 `gloo` tests (CPU) of the multi-GPU plumbing in contrast_you_b200/distributed.py.

The CUDA kernels cannot run here, so each rank evaluates ITS shard with the CPU oracle (test infrastructure) and the
product's exchange code (gather_rank_major / local_view_major / gather_rows_ / make_joint_reduce) has to turn the
per-rank pieces into exactly the single-process result."""
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
        # what ShardedSupConLoss does: canonical labels and sort order are LOCAL, then three in-place all-gathers
        lab_loc = cyd.local_view_major(lab[sl].to(torch.float32), _canon_cpu)  # [2 n_loc] canonical, view-major
        order = torch.argsort(lab_loc)
        z_all = cyd.gather_rank_major(local[order])                            # [G * 2 n_loc, d], rank-major, blocks sorted
        rb, re = cyd.row_range(n_loc)
        assert (rb, re) == (rank * 2 * n_loc, (rank + 1) * 2 * n_loc)
        N = 2 * n
        labels = torch.zeros(N, dtype=torch.int32)
        labels[rb:re] = lab_loc[order]
        cyd.gather_rows_(labels)
        # every rank's block holds ITS samples' labels, sorted: compare with a single-process construction
        for r in range(WORLD):
            want = _canon_cpu(lab[r * n_loc:(r + 1) * n_loc].to(torch.float32), n_loc).sort().values
            assert torch.equal(labels[r * 2 * n_loc:(r + 1) * 2 * n_loc], want)
        # this rank's strip with the oracle's closed form: xstat rows of the owned rows only
        Z = z_all.detach().numpy()
        lb = labels.numpy()
        S = Z[rb:re] @ Z.T / t
        m = 1.0 / t
        rows = np.arange(rb, re)
        P = (lb[rb:re, None] == lb[None, :]).astype(np.float64); P[rows - rb, rows] = 0
        E = np.exp(S - m); E[rows - rb, rows] = 0
        D, c = E.sum(1), P.sum(1)
        xstat = torch.zeros(N, 4, dtype=torch.float64)                          # CY_XS_*: logden, 1/c, coef, loss term
        xstat[rb:re, 0] = torch.from_numpy(np.log(D + 1e-16))
        xstat[rb:re, 1] = torch.from_numpy(1 / c)
        xstat[rb:re, 2] = torch.from_numpy(1 / (D + 1e-16))
        xstat[rb:re, 3] = torch.from_numpy(-((P * (S - m)).sum(1) / c - np.log(D + 1e-16)))
        cyd.gather_rows_(xstat)                                                # the product's ONE statistics exchange
        loss = xstat[:, 3].sum().item() / N                                    # what cy_infonce_loss reduces on every rank
        coef, invc = xstat[:, 2].numpy(), xstat[:, 1].numpy()
        W = E * (coef[rb:re, None] + coef[None, :]) - P * (invc[rb:re, None] + invc[None, :])
        dz_all = torch.zeros(N, d, dtype=torch.float64)
        dz_all[rb:re] = torch.from_numpy(W @ Z / (t * N))
        z_all.backward(dz_all)                                                 # through _GatherRows.backward + the local un-sort
        # single-process truth: the literal oracle on the concatenated batch
        ref = OC.supcon(f1.numpy(), f2.numpy(), target=lab.tolist(), t=t)
        assert abs(loss - ref["loss"]) < 1e-12 * abs(ref["loss"]), (loss, ref["loss"])
        got = local.grad.numpy()
        want = np.concatenate([ref["grad_f1"][sl], ref["grad_f2"][sl]])
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()

        # ---------------- batch-sharded IIC: all-reduce of the raw joint before the epilogue
        g = torch.Generator().manual_seed(8)
        x = torch.randn(2 * WORLD, 5, 9, 11, generator=g, dtype=torch.float64).softmax(1).numpy()
        y = torch.randn(2 * WORLD, 5, 9, 11, generator=g, dtype=torch.float64).softmax(1).numpy()
        bs = slice(2 * rank, 2 * rank + 2)
        J_loc = torch.from_numpy(OM.raw_joint_2d(x[bs], y[bs], 1))
        # the callback protocol of IIDSegmentationLoss (collective form; the peer-memory form needs CUDA)
        J, n_slots, npx = cyd.make_joint_reduce(exchange="nccl")(lambda out: out.copy_(J_loc), tuple(J_loc.shape), float(2 * 9 * 11),
                                                                   torch.device("cpu"))
        assert npx == 2 * WORLD * 9 * 11 and n_slots == 1 and J.dtype == torch.float64
        np.testing.assert_allclose(J.numpy(), OM.raw_joint_2d(x, y, 1), rtol=1e-13)
        # the sub-head stack (forward_heads under sharding): S partial joints travel as ONE [S, K,K,T,T] array and the mean of the
        # per-head losses on the reduced joints equals the single-process sub-head mean on the gathered batch
        S = 3
        xs = [torch.randn(2 * WORLD, 5, 9, 11, generator=g, dtype=torch.float64).softmax(1).numpy() for _ in range(S)]
        ys = [torch.randn(2 * WORLD, 5, 9, 11, generator=g, dtype=torch.float64).softmax(1).numpy() for _ in range(S)]
        stack_loc = torch.from_numpy(np.stack([OM.raw_joint_2d(a[bs], b[bs], 1) for a, b in zip(xs, ys)]))
        Js, n_slots, npx = cyd.make_joint_reduce(exchange="nccl")(lambda out: out.copy_(stack_loc), tuple(stack_loc.shape),
                                                                    float(2 * 9 * 11), torch.device("cpu"))
        assert tuple(Js.shape) == (S, 5, 5, 3, 3) and n_slots == 1 and npx == 2 * WORLD * 9 * 11
        got = 0.0
        for s_ in range(S):
            P, _aux = OM.joint_epilogue(Js[s_].numpy(), padding=1, symmetric=False, n_pixels=npx)
            got += OM.mi_loss_and_grad_wrt_pij(P, 1.0, 1e-5)[0] / S
        want = sum(OM.iid_segmentation_loss(a, b, padding=1)["loss"] for a, b in zip(xs, ys)) / S
        assert abs(got - want) < 1e-12 * abs(want)
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
