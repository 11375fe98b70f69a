#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference tree is copied to a scratch directory (its ``contrastyou/__init__.py`` mkdirs at import
time and /root/reference is read-only) and imported from there.  ``contrastyou.losses.discreteMI``
pulls in plotting / medical-imaging packages that are not installed; they are replaced by empty shim
modules *before* import.  No reference arithmetic is restated here: every number
written to the fixtures is produced by the reference's own code (float64 CPU, torch autograd).

Fixtures (all small, a few hundred KB in total):
  supcon_*.npz   inputs f1,f2 (+labels/mask) -> loss, grad_f1, grad_f2, pos_mask, neg_mask, sim_exp, sim_logits
  selfpaced_*.npz  same + downgrade_ratio, sp_mask
  iic_*.npz      inputs x,y (+mask) -> loss, grad_x, grad_y, joint (get_joint_matrix)
  iid_*.npz      IIDLoss 3-tuple + grads
  labels.json    label generators (LabelEncoder ranks etc.)
  regions.json   region_extractor coordinates for given (h, w, seed)
  heads_*.npz    projector head outputs for a fixed state_dict (key layout pin)
"""
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CY_REFERENCE", "/root/reference")


def _import_reference():
    scratch = tempfile.mkdtemp(prefix="cy_ref_")
    for sub in ("contrastyou", "semi_seg", "script"):
        shutil.copytree(os.path.join(REF, sub), os.path.join(scratch, sub))
    sys.path.insert(0, scratch)
    os.environ.setdefault("LOGURU_LEVEL", "ERROR")

    def shim(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    shim("termcolor", colored=lambda s, *a, **k: s)
    mpl = shim("matplotlib", use=lambda *a, **k: None, get_backend=lambda: "agg")
    mpl.pyplot = shim("matplotlib.pyplot", switch_backend=lambda *a, **k: None)
    medpy = shim("medpy")
    medpy.metric = shim("medpy.metric", assd=None)
    medpy.metric.binary = shim("medpy.metric.binary", __surface_distances=None)

    import contrastyou  # noqa  (creates .data/ runs/ inside the scratch copy)
    from contrastyou.losses.kl import Entropy
    # pre-seed semi_seg.hooks.midl so discreteMI.py:17 does not import the whole hook tree
    pkg = shim("semi_seg"); pkg.__path__ = [os.path.join(scratch, "semi_seg")]
    hooks = shim("semi_seg.hooks"); hooks.__path__ = [os.path.join(scratch, "semi_seg", "hooks")]
    shim("semi_seg.hooks.midl", entropy_criterion=Entropy(reduction="none", eps=1e-8))
    return scratch


def _np(t):
    return t.detach().cpu().numpy()


def supcon_case(name, n, d, seed, *, labels=None, mask=None, exclude=False, dtype=torch.float64, t=0.07,
                selfpaced=None):
    from contrastyou.losses.contrastive import SupConLoss1, SelfPacedSupConLoss
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    f1 = F.normalize(torch.randn(n, d, dtype=dtype, generator=g), dim=1).requires_grad_()
    f2 = F.normalize(torch.randn(n, d, dtype=dtype, generator=g), dim=1).requires_grad_()
    kw = {}
    out = {}
    if labels == "randint5":
        lab = torch.randint(0, 5, (n,), generator=g).tolist()
    elif labels == "partition3":
        lab = torch.randint(0, 3, (n,), generator=g).tolist()
    elif labels == "self":
        lab = list(range(n))
    elif labels == "big":   # float32 collision quirk: 2**24 and 2**24+1 compare equal (contrastive.py:40)
        lab = [2 ** 24 + (i % 4) for i in range(n)]
    else:
        lab = labels
    if lab is not None:
        kw["target"] = lab
        out["labels"] = np.asarray(lab, dtype=np.int64)
    if mask is not None:
        if mask == "random3":
            m = torch.randint(0, 3, (n, n), generator=g).to(dtype)
            m[torch.arange(n), torch.arange(n)] = 1  # every row keeps at least one positive (its twin view)
        kw["mask"] = m
        out["mask"] = _np(m)
    if selfpaced is None:
        crit = SupConLoss1(temperature=t, exclude_other_pos=exclude)
    else:
        crit = SelfPacedSupConLoss(temperature=t, weight_update=selfpaced["mode"],
                                   correct_grad=selfpaced.get("correct_grad", False))
        crit.set_gamma(selfpaced["gamma"])
    loss = crit(f1, f2, **kw)
    loss.backward()
    out.update(f1=_np(f1), f2=_np(f2), loss=_np(loss), grad_f1=_np(f1.grad), grad_f2=_np(f2.grad),
               pos_mask=_np(crit.pos_mask).astype(np.uint8), neg_mask=_np(crit.neg_mask).astype(np.uint8),
               temperature=np.float64(t), exclude=np.bool_(exclude))
    if n <= 16:
        out.update(sim_exp=_np(crit.sim_exp), sim_logits=_np(crit.sim_logits))
    if selfpaced is not None:
        out.update(downgrade_ratio=np.float64(crit.downgrade_ratio), sp_mask=_np(crit.sp_mask),
                   gamma=np.float64(selfpaced["gamma"]), mode=np.str_(selfpaced["mode"]),
                   correct_grad=np.bool_(selfpaced.get("correct_grad", False)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: loss={loss.item()!r}")
    return loss.item()


def iic_case(name, B, K, H, W, seed, *, padding, symmetric=False, lamda=1.0, eps=1e-5, with_mask=False,
             scale=1.0):
    from contrastyou.losses.discreteMI import IIDSegmentationLoss
    g = torch.Generator().manual_seed(seed)
    lx = (scale * torch.randn(B, K, H, W, dtype=torch.float64, generator=g)).requires_grad_()
    ly = (scale * torch.randn(B, K, H, W, dtype=torch.float64, generator=g)).requires_grad_()
    x, y = lx.softmax(1), ly.softmax(1)
    x.retain_grad(); y.retain_grad()
    x_in, y_in = _np(x).copy(), _np(y).copy()
    out = {}
    kw = {}
    if with_mask:
        m = (torch.rand(B, 1, H, W, generator=g) > 0.3).to(torch.float64)
        kw["mask"] = m
        out["mask"] = _np(m)
        # the reference multiplies in place (discreteMI.py:142-144): feed non-leaf copies so autograd allows it
        xin, yin = x * 1.0, y * 1.0
        xin.retain_grad(); yin.retain_grad()
    else:
        xin, yin = x, y
    crit = IIDSegmentationLoss(lamda=lamda, padding=padding, eps=eps, symmetric=symmetric)
    loss = crit(xin, yin, **kw)
    loss.backward()
    out.update(x=x_in, y=y_in, loss=_np(loss), grad_x=_np(x.grad), grad_y=_np(y.grad),
               joint=crit.get_joint_matrix(), padding=np.int64(padding), symmetric=np.bool_(symmetric),
               lamda=np.float64(lamda), eps=np.float64(eps))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: loss={loss.item()!r} sum|gx|={x.grad.abs().sum().item()!r}")
    return loss.item()


def logits_case(name, S, B, K, H, W, seed, *, T, padding, symmetric=False):
    """the cluster-head tail + the discrete-MI hook's criterion call, run by the reference itself: ``SoftmaxWithT(1, T)``
    (projectors/nn.py:36-44) on S pairs of logits, then ``sum(criterion(x1, x2) ...) / S`` (semi_seg/hooks/discretemi.py:111);
    gradients w.r.t. the LOGITS.  Pins IIDSegmentationLoss.forward_heads(..., logits_T=T)."""
    from contrastyou.losses.discreteMI import IIDSegmentationLoss
    from contrastyou.projectors.nn import SoftmaxWithT
    g = torch.Generator().manual_seed(seed)
    lx = [(2 * torch.randn(B, K, H, W, dtype=torch.float64, generator=g)).requires_grad_() for _ in range(S)]
    ly = [(2 * torch.randn(B, K, H, W, dtype=torch.float64, generator=g)).requires_grad_() for _ in range(S)]
    tail = SoftmaxWithT(1, T=T)
    crit = IIDSegmentationLoss(padding=padding, symmetric=symmetric)
    # the tail divides in place (nn.py:43): feed non-leaf copies, as the 1x1 conv output is in the reference
    loss = sum(crit(tail(a * 1.0), tail(b * 1.0)) for a, b in zip(lx, ly)) / S
    loss.backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), logits_x=np.stack([_np(t) for t in lx]),
                        logits_y=np.stack([_np(t) for t in ly]), loss=_np(loss), grad_x=np.stack([_np(t.grad) for t in lx]),
                        grad_y=np.stack([_np(t.grad) for t in ly]), T=np.float64(T), padding=np.int64(padding),
                        symmetric=np.bool_(symmetric))
    print(f"{name}: loss={loss.item()!r} sum|glx0|={lx[0].grad.abs().sum().item()!r}")
    return loss.item()


def iid_case(name, bn, K, seed, lamb=1.0):
    from contrastyou.losses.discreteMI import IIDLoss
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(bn, K, dtype=torch.float64, generator=g).softmax(1).requires_grad_()
    y = torch.randn(bn, K, dtype=torch.float64, generator=g).softmax(1).requires_grad_()
    loss, loss_no_lamb, pij = IIDLoss(lamb=lamb)(x, y)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=_np(x), y=_np(y), loss=_np(loss),
                        loss_no_lamb=_np(loss_no_lamb), p_i_j=_np(pij), grad_x=_np(x.grad), grad_y=_np(y.grad),
                        lamb=np.float64(lamb))
    print(f"{name}: loss={loss.item()!r}")


def sibling_cases():
    """losses that reuse the IIC joint with a different epilogue: RedundancyCriterion
    (redundancy_reduction.py:12-33), PUISegLoss (pica_loss.py:43-80), IMSATLoss / IMSATDynamicWeight / imsat_loss
    (discreteMI.py:20-87, 275-297)."""
    from contrastyou.losses.redundancy_reduction import RedundancyCriterion
    from contrastyou.losses.pica_loss import PUISegLoss
    from contrastyou.losses.discreteMI import IMSATLoss, IMSATDynamicWeight

    def maps(B, K, H, W, seed):
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, K, H, W, dtype=torch.float64, generator=g).softmax(1).requires_grad_()
        y = torch.randn(B, K, H, W, dtype=torch.float64, generator=g).softmax(1).requires_grad_()
        return x, y

    for name, sym, alpha, lamda in (("redundancy_sym", True, 0.3, 1.0), ("redundancy_asym", False, 0.7, 2.0)):
        x, y = maps(2, 6, 16, 16, 40)
        crit = RedundancyCriterion(symmetric=sym, lamda=lamda, alpha=alpha)
        loss = crit(x, y)
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=_np(x), y=_np(y), loss=_np(loss), grad_x=_np(x.grad),
                            grad_y=_np(y.grad), joint=crit.get_joint_matrix(), symmetric=np.bool_(sym), alpha=np.float64(alpha),
                            lamda=np.float64(lamda))
        print(f"{name}: loss={loss.item()!r}")
    for name, pad, lamda in (("puiseg_pad1", 1, 2.0), ("puiseg_pad3", 3, 0.5)):
        x, y = maps(2, 5, 14, 18, 41)
        loss = PUISegLoss(lamda=lamda, padding=pad)(x, y)
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=_np(x), y=_np(y), loss=_np(loss), grad_x=_np(x.grad),
                            grad_y=_np(y.grad), padding=np.int64(pad), lamda=np.float64(lamda))
        print(f"{name}: loss={loss.item()!r}")
    g = torch.Generator().manual_seed(42)
    x = torch.randn(16, 5, dtype=torch.float64, generator=g).softmax(1).requires_grad_()
    y = torch.randn(16, 5, dtype=torch.float64, generator=g).softmax(1).requires_grad_()
    l2 = IMSATLoss(lamda=1.5)(x, y)
    l2.backward()
    gx2, gy2 = _np(x.grad).copy(), _np(y.grad).copy()
    x.grad = None
    l1 = IMSATLoss(lamda=1.5)(x)
    l1.backward()
    gx1 = _np(x.grad).copy()
    x.grad = None
    dyn = IMSATDynamicWeight(lamda=0.8)
    ld = dyn(x)
    ld.backward()
    np.savez_compressed(os.path.join(HERE, "imsat.npz"), x=_np(x), y=_np(y), loss_pair=_np(l2), grad_x_pair=gx2, grad_y_pair=gy2,
                        loss_single=_np(l1), grad_x_single=gx1, loss_dynamic=_np(ld), grad_x_dynamic=_np(x.grad),
                        dynamic_weight_after=_np(dyn.dynamic_weight), lamda=np.float64(1.5), lamda_dynamic=np.float64(0.8))
    print(f"imsat: pair={l2.item()!r} single={l1.item()!r} dynamic={ld.item()!r}")


def _exec_defs(relpath, names, ns):
    """exec selected top-level defs of a reference source file (its module cannot be imported here because of
    absent third-party packages); the code object is the reference's own, nothing is restated."""
    import ast
    tree = ast.parse(open(os.path.join(REF, relpath)).read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(body) == len(names), (relpath, names)
    exec(compile(ast.Module(body=body, type_ignores=[]), relpath, "exec"), ns)
    return [ns[n] for n in names]


def label_cases():
    from typing import List
    from sklearn.preprocessing import LabelEncoder
    PartitionLabelGenerator, PatientLabelGenerator, ACDCCycleGenerator, SIMCLRGenerator = _exec_defs(
        "semi_seg/epochers/helper.py",
        ["PartitionLabelGenerator", "PatientLabelGenerator", "ACDCCycleGenerator", "SIMCLRGenerator"],
        {"List": List, "LabelEncoder": LabelEncoder})
    partition = ["1", "0", "2", "2", "0", "1", "10", "9"]
    patient = ["patient003", "patient001", "patient003", "patient100", "patient020", "patient001", "a", "B"]
    experiment = ["00", "01", "00", "00", "01", "01", "00", "02"]
    cases = {
        "partition": {"in": partition, "out": PartitionLabelGenerator()(partition_list=partition)},
        "patient": {"in": patient, "out": PatientLabelGenerator()(patient_list=patient)},
        "cycle": {"in": experiment, "out": ACDCCycleGenerator()(experiment_list=experiment)},
        "self": {"in": partition, "out": SIMCLRGenerator()(partition_list=partition)},
    }
    with open(os.path.join(HERE, "labels.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("labels:", {k: v["out"] for k, v in cases.items()})


def head_cases():
    """Pin the projector heads: state_dict key layout + outputs for seeded weights (heads.py:81-200)."""
    from contrastyou.projectors.heads import (ProjectionHead, DenseProjectionHead, ClusterHead, DenseClusterHead,
                                              CrossCorrelationProjector)
    torch.manual_seed(7)
    feats = torch.randn(4, 16, 8, 8, dtype=torch.float64)
    specs = {
        "proj_mlp": (ProjectionHead, dict(input_dim=16, hidden_dim=24, output_dim=12, head_type="mlp", normalize=True)),
        "proj_linear": (ProjectionHead, dict(input_dim=16, output_dim=12, head_type="linear", normalize=False,
                                             pool_name="adaptive_max")),
        "dense_mlp": (DenseProjectionHead, dict(input_dim=16, hidden_dim=24, output_dim=12, head_type="mlp",
                                                normalize=True, spatial_size=(4, 4))),
        "dense_linear_id": (DenseProjectionHead, dict(input_dim=16, output_dim=12, head_type="linear",
                                                      normalize=True, pool_name="identical")),
        "cluster_linear": (ClusterHead, dict(input_dim=16, num_clusters=5, num_subheads=3, head_type="linear", T=2,
                                             normalize=False)),
        "cluster_mlp": (ClusterHead, dict(input_dim=16, num_clusters=5, num_subheads=2, head_type="mlp", T=1,
                                          normalize=True)),
        "dense_cluster_linear": (DenseClusterHead, dict(input_dim=16, num_clusters=6, num_subheads=2,
                                                        head_type="linear", T=1, normalize=False)),
        "dense_cluster_mlp": (DenseClusterHead, dict(input_dim=16, num_clusters=6, hidden_dim=10, num_subheads=2,
                                                     head_type="mlp", T=0.5, normalize=True)),
        "cc_projector": (CrossCorrelationProjector, dict(input_dim=16, num_clusters=7, head_type="mlp",
                                                         normalize=False, T=1.0, num_subheads=2, hidden_dim=9)),
    }
    meta = {}
    for name, (cls, kw) in specs.items():
        torch.manual_seed(11)
        head = cls(**kw).double()
        sd = {k: _np(v) for k, v in head.state_dict().items()}
        out = head(feats.clone())
        outs = out if isinstance(out, list) else [out]
        arrays = {"feats": _np(feats)}
        arrays.update({"sd::" + k: v for k, v in sd.items()})
        arrays.update({f"out{i}": _np(o) for i, o in enumerate(outs)})
        np.savez_compressed(os.path.join(HERE, f"heads_{name}.npz"), **arrays)
        meta[name] = {"class": cls.__name__, "kwargs": {k: (list(v) if isinstance(v, tuple) else v)
                                                        for k, v in kw.items()},
                      "keys": list(sd.keys()), "is_list": isinstance(out, list)}
    with open(os.path.join(HERE, "heads.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("heads:", list(meta))


def region_cases():
    """region_extractor coordinates (semi_seg/hooks/infonce.py:31-46) — the full hook module cannot be imported
    (matplotlib, PIL writer stack), so the function object is exec'd from its own source lines."""
    from contrastyou.utils.utils import fix_all_seed_for_transforms
    region_extractor, = _exec_defs("semi_seg/hooks/infonce.py", ["region_extractor"],
                                   {"np": np, "torch": torch,
                                    "fix_all_seed_for_transforms": fix_all_seed_for_transforms})
    cases = []
    for (b, c, h, w, seed, pn) in [(3, 4, 10, 10, 1, 5), (2, 3, 20, 12, 12345, 5), (4, 2, 16, 16, 7, 8)]:
        # feature value encodes its own (h, w) coordinate so the gather order can be read back
        fm = torch.zeros(b, c, h, w)
        hh, ww = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        fm[:, 0] = hh.float()
        fm[:, 1] = ww.float()
        out = region_extractor(fm, point_nums=pn, seed=seed)
        coords = out[:, :2].long().tolist()
        cases.append({"b": b, "c": c, "h": h, "w": w, "seed": seed, "point_nums": pn, "coords": coords})
    with open(os.path.join(HERE, "regions.json"), "w") as f:
        json.dump(cases, f)
    print("regions:", [len(c["coords"]) for c in cases])


def logits_cases():
    logits_case("logits_iic_pad1_T05", 2, 2, 10, 12, 16, 40, T=0.5, padding=1)
    logits_case("logits_iic_pad1_sym_T2", 2, 2, 6, 12, 16, 41, T=2.0, padding=1, symmetric=True)
    logits_case("logits_iic_pad0_T1", 2, 3, 5, 8, 8, 42, T=1.0, padding=0)


def main():
    scratch = _import_reference()
    known = {}
    if "--only-logits" in sys.argv:      # added after the other fixtures: regenerate these alone
        try:
            logits_cases()
        finally:
            sys.path.remove(scratch)
            shutil.rmtree(scratch, ignore_errors=True)
        return
    try:
        # --- SupConLoss1 (contrastive.py:23-100)
        # survey-time known answers: torch.manual_seed(0) global stream
        import torch.nn.functional as F
        from contrastyou.losses.contrastive import SupConLoss1
        torch.manual_seed(0)
        f1 = F.normalize(torch.randn(64, 256, dtype=torch.float64))
        f2 = F.normalize(torch.randn(64, 256, dtype=torch.float64))
        labels = torch.randint(0, 5, (64,)).tolist()
        known["supcon_survey"] = SupConLoss1()(f1, f2, target=labels).item()
        known["simclr_survey"] = SupConLoss1()(f1, f2).item()
        known["exclude_survey"] = SupConLoss1(exclude_other_pos=True)(f1, f2, target=labels).item()
        np.savez_compressed(os.path.join(HERE, "supcon_survey.npz"), f1=_np(f1), f2=_np(f2),
                            labels=np.asarray(labels), **{k: np.float64(v) for k, v in known.items()})
        assert abs(known["supcon_survey"] - 5.235924850353655) < 1e-12, known

        supcon_case("supcon_labels5", 24, 32, 1, labels="randint5")
        supcon_case("supcon_partition3", 64, 64, 2, labels="partition3")
        supcon_case("supcon_simclr", 16, 32, 3)
        supcon_case("supcon_self", 16, 32, 3, labels="self")
        supcon_case("supcon_exclude", 24, 32, 4, labels="randint5", exclude=True)
        supcon_case("supcon_mask3", 12, 16, 5, mask="random3")
        supcon_case("supcon_mask3_exclude", 12, 16, 5, mask="random3", exclude=True)
        supcon_case("supcon_biglabels", 8, 16, 6, labels="big")
        supcon_case("supcon_tiny", 2, 8, 7, labels=[0, 0])
        supcon_case("supcon_temp05", 16, 32, 8, labels="randint5", t=0.5)
        supcon_case("supcon_d256", 32, 256, 9, labels="partition3")
        # --- SelfPacedSupConLoss (contrastive.py:103-212)
        supcon_case("selfpaced_soft6", 24, 32, 10, labels="randint5", selfpaced=dict(mode="soft", gamma=6.0))
        supcon_case("selfpaced_hard5", 24, 32, 10, labels="randint5", selfpaced=dict(mode="hard", gamma=5.0))
        supcon_case("selfpaced_hard5_cg", 24, 32, 10, labels="randint5",
                    selfpaced=dict(mode="hard", gamma=5.0, correct_grad=True))
        supcon_case("selfpaced_soft_inf", 24, 32, 10, labels="randint5", selfpaced=dict(mode="soft", gamma=1e10))
        supcon_case("selfpaced_default", 16, 32, 11, selfpaced=dict(mode="hard", gamma=1e6))
        # --- IIDSegmentationLoss (discreteMI.py:127-170, 225-261)
        from contrastyou.losses.discreteMI import IIDSegmentationLoss
        torch.manual_seed(0)
        x = torch.randn(4, 10, 32, 32, dtype=torch.float64).softmax(1)
        y = torch.randn(4, 10, 32, 32, dtype=torch.float64).softmax(1)
        known["iic_pad1_survey"] = IIDSegmentationLoss(padding=1)(x, y).item()
        known["iic_pad0_survey"] = IIDSegmentationLoss(padding=0)(x, y).item()
        assert abs(known["iic_pad1_survey"] - (-0.2478573718202518)) < 1e-12, known
        np.savez_compressed(os.path.join(HERE, "iic_survey.npz"), x=_np(x), y=_np(y),
                            loss_pad1=np.float64(known["iic_pad1_survey"]),
                            loss_pad0=np.float64(known["iic_pad0_survey"]))
        iic_case("iic_pad1", 3, 10, 20, 24, 20, padding=1)
        iic_case("iic_pad1_sym", 3, 10, 20, 24, 20, padding=1, symmetric=True)
        iic_case("iic_pad0", 3, 10, 20, 24, 21, padding=0)
        iic_case("iic_pad0_sym", 3, 10, 20, 24, 21, padding=0, symmetric=True)
        iic_case("iic_pad2_lam15", 2, 5, 17, 13, 22, padding=2, lamda=1.5, symmetric=True)
        iic_case("iic_pad1_mask", 2, 6, 16, 16, 23, padding=1, with_mask=True)
        iic_case("iic_pad3_k20", 2, 20, 12, 12, 24, padding=3, scale=2.0)
        iic_case("iic_pad1_sharp", 2, 10, 16, 16, 25, padding=1, scale=6.0)
        # --- IIDLoss (discreteMI.py:90-124, 201-222)
        iid_case("iid_k20", 18, 20, 30)
        iid_case("iid_k5_lam2", 7, 5, 31, lamb=2.0)
        logits_cases()
        sibling_cases()
        label_cases()
        region_cases()
        head_cases()
        with open(os.path.join(HERE, "known_answers.json"), "w") as f:
            json.dump(known, f, indent=1)
    finally:
        sys.path.remove(scratch)
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
