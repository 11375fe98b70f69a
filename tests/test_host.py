"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares (no compute calls), the
host-side mirrors (labels, point sampling, projector heads) match the golden fixtures, and the product path refuses
to run without CUDA instead of falling back."""
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden

import contrast_you_b200  # noqa: F401
from contrast_you_b200 import _lib, labels as lab, sampling
from contrast_you_b200 import projectors as P
from contrast_you_b200.losses import SupConLoss1, SelfPacedSupConLoss, IIDSegmentationLoss, IIDLoss


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "contrastyou_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cy_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_header_symbol():
    names = _declared_symbols()
    assert len(names) >= 14, names
    handle = _lib.load()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in contrastyou_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert handle.cy_abi_version() == _lib.CY_ABI_VERSION


def test_argument_validation_needs_no_gpu():
    handle = _lib.load()
    # null pointers / bad sizes are rejected before anything touches the device
    assert handle.cy_infonce_fwd(None, 0, 4, 8, 8, None, None, 0, 4, 1.0, 0, 0, None, None, None, 0, None) == -1
    assert b"null" in handle.cy_last_error()
    assert handle.cy_iic_joint(None, None, 0, 1, 2, 3, 3, 1, None, None, 0, None) == -1
    assert handle.cy_iic_workspace_bytes(0, 1, 1, 1, 0) == 0


def test_no_cpu_fallback():
    f = torch.nn.functional.normalize(torch.randn(4, 8), dim=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        SupConLoss1()(f, f, target=[0, 0, 1, 1])
    x = torch.randn(2, 3, 4, 4).softmax(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        IIDSegmentationLoss(padding=1)(x, x)


def test_modules_own_no_state():
    for m in (SupConLoss1(), SelfPacedSupConLoss(weight_update="soft", correct_grad=True), IIDSegmentationLoss(padding=1),
              IIDLoss()):
        assert len(m.state_dict()) == 0 and len(list(m.parameters())) == 0 and len(list(m.buffers())) == 0


def test_module_surface_matches_reference():
    c = SelfPacedSupConLoss()
    assert c.age_param == 1e6
    c.set_gamma(3)
    assert c.age_param == 3.0 and isinstance(c.age_param, float)
    m = IIDSegmentationLoss(lamda=1.5, padding=2, symmetric=True)
    assert (m.lamda, m.padding, m.symmetric, m._eps) == (1.5, 2, True, 1e-5)
    with pytest.raises(RuntimeError):
        m.get_joint_matrix()
    assert IIDLoss(lamb=2).lamb == 2.0


def test_label_generators_golden():
    cases = json.load(open(os.path.join(GOLDEN, "labels.json")))
    assert lab.PartitionLabelGenerator()(partition_list=cases["partition"]["in"]) == cases["partition"]["out"]
    assert lab.PatientLabelGenerator()(patient_list=cases["patient"]["in"]) == cases["patient"]["out"]
    assert lab.ACDCCycleGenerator()(experiment_list=cases["cycle"]["in"]) == cases["cycle"]["out"]
    assert lab.SIMCLRGenerator()(partition_list=cases["self"]["in"]) == cases["self"]["out"]
    groups = [f"{p}_{e}" for p, e in zip(cases["patient"]["in"], cases["cycle"]["in"])]
    part = cases["partition"]["in"]
    assert lab.get_label("patient", "acdc", part, groups) == cases["patient"]["out"]
    assert lab.get_label("cycle", "acdc", part, groups) == cases["cycle"]["out"]
    assert lab.get_label("partition", "prostate_md", part, groups) == cases["partition"]["out"]
    assert lab.get_label("self", "spleen", part, groups) == cases["self"]["out"]
    with pytest.raises(NotImplementedError):
        lab.get_label("cycle", "spleen", part, groups)
    with pytest.raises(NotImplementedError):
        lab.get_label("partition", "imagenet", part, groups)


def test_region_extractor_golden():
    for c in json.load(open(os.path.join(GOLDEN, "regions.json"))):
        fm = torch.zeros(c["b"], c["c"], c["h"], c["w"])
        hh, ww = torch.meshgrid(torch.arange(c["h"]), torch.arange(c["w"]), indexing="ij")
        fm[:, 0], fm[:, 1] = hh.float(), ww.float()
        state = np.random.get_state()[1].copy()
        out = sampling.region_extractor(fm, point_nums=c["point_nums"], seed=c["seed"])
        assert out.shape == (c["b"] * c["point_nums"], c["c"])
        assert out[:, :2].long().tolist() == c["coords"]
        assert (np.random.get_state()[1] == state).all()      # global numpy RNG untouched, like the reference's context


@pytest.mark.parametrize("name", sorted(json.load(open(os.path.join(GOLDEN, "heads.json")))))
def test_projector_heads_golden(name):
    meta = json.load(open(os.path.join(GOLDEN, "heads.json")))[name]
    g = load_golden("heads_" + name)
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in meta["kwargs"].items()}
    head = getattr(P, meta["class"])(**kw).double()
    assert list(head.state_dict().keys()) == meta["keys"]           # checkpoint key layout
    head.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=True)
    out = head(torch.from_numpy(g["feats"]).clone())
    assert isinstance(out, list) == meta["is_list"]
    outs = out if isinstance(out, list) else [out]
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.detach().numpy(), g[f"out{i}"], rtol=1e-10, atol=1e-12)


def test_imsat_losses_match_reference_fixture():
    """IMSATLoss / IMSATDynamicWeight (discreteMI.py:20-87, :275-297): pure host-side mirrors, checked on CPU in float64"""
    import numpy as np
    import torch
    from conftest import load_golden
    from contrast_you_b200.losses import IMSATLoss, IMSATDynamicWeight
    g = load_golden("imsat")
    x = torch.tensor(g["x"], requires_grad=True)
    y = torch.tensor(g["y"], requires_grad=True)
    loss = IMSATLoss(lamda=float(g["lamda"]))(x, y)
    loss.backward()
    assert abs(loss.item() - float(g["loss_pair"])) < 1e-12
    np.testing.assert_allclose(x.grad.numpy(), g["grad_x_pair"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(y.grad.numpy(), g["grad_y_pair"], rtol=1e-10, atol=1e-14)
    x.grad = None
    loss = IMSATLoss(lamda=float(g["lamda"]))(x)
    loss.backward()
    assert abs(loss.item() - float(g["loss_single"])) < 1e-12
    np.testing.assert_allclose(x.grad.numpy(), g["grad_x_single"], rtol=1e-10, atol=1e-14)
    x.grad = None
    dyn = IMSATDynamicWeight(lamda=float(g["lamda_dynamic"]))
    loss = dyn(x)
    loss.backward()
    assert abs(loss.item() - float(g["loss_dynamic"])) < 1e-12
    np.testing.assert_allclose(x.grad.numpy(), g["grad_x_dynamic"], rtol=1e-10, atol=1e-14)
    assert abs(float(dyn.dynamic_weight) - float(g["dynamic_weight_after"])) < 1e-12
    assert "dynamic_weight" in dyn.state_dict()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) prints ONE JSON line with the agreed keys,
    the same metric / unit / config as our arm, and needs neither a GPU nor /root/reference."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "InfoNCE fwd+bwd pairs/s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["workload"].startswith("cfg4")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cluster_head_skip_softmax_returns_the_logits_of_the_same_module_tree():
    """DenseClusterHead(features, skip_softmax=True): logits such that softmax(logits / T) == the stock forward; state_dict
    keys unchanged (the SoftmaxWithT tail stays a member, it is bypassed at call time — SURVEY.md 8b)"""
    import torch
    from contrast_you_b200.projectors import DenseClusterHead
    torch.manual_seed(0)
    head = DenseClusterHead(input_dim=8, num_clusters=5, num_subheads=3, T=2.0, head_type="mlp", hidden_dim=16)
    feats = torch.randn(2, 8, 6, 6)
    logits = head(feats, skip_softmax=True)
    probs = head(feats)
    assert len(logits) == 3 and head.temperature == 2.0
    for lg, pr in zip(logits, probs):
        assert torch.allclose(torch.softmax(lg / 2.0, 1), pr, atol=1e-6)
    assert sorted(head.state_dict()) == sorted(f"_headers.{s}.{i}.{w}" for s in range(3) for i in (0, 2) for w in ("weight", "bias"))


def test_fp32_reciprocal_index_split_is_exact_below_2_22():
    """cy_iic_epilogue splits flat indices by the runtime K / T*T / K*K as int((n + 0.5f) * (1.0f / d)) (csrc/iic.cu): exact for
    every n < 2^22 — the bound its host wrapper enforces — restated with numpy float32 for the divisors that can occur"""
    n = np.arange(1 << 22, dtype=np.int64)
    nf = n.astype(np.float32) + np.float32(0.5)
    for d in list(range(1, 65)) + [100, 121, 225, 256, 400, 49, 25, 9, 4096, 65536]:
        q = (nf * (np.float32(1.0) / np.float32(d))).astype(np.int32)
        assert np.array_equal(q, (n // d).astype(np.int32)), d


def test_header_is_plain_c():
    """the drop-in boundary is a C ABI: include/contrastyou_b200.h must compile as C99 (and as C++) on its own — plain pointers and
    sizes, no torch / CUDA types in the signatures"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "contrastyou_b200.h")
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    for args in (["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c"], ["g++", "-std=c++17", "-fsyntax-only", "-x", "c++"]):
        out = subprocess.run(args + [hdr], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)      # declarations only, comments stripped
    assert "torch" not in code.lower() and "cudaStream_t" not in code and "at::" not in code
