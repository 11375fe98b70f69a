"""The oracle against the golden vectors the reference itself produced (tests/golden/make_golden.py).

CPU only.  These tests are what "pins" the oracle: every loss / gradient / mask in the fixtures comes from
running the unmodified reference modules (float64, torch autograd)."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import contrastive_np as C
from oracle import discrete_mi_np as M
from oracle import labels_np as Lb
from oracle import c_oracle

SUPCON = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "supcon_*.npz"))
                if "survey" not in p)
SELFPACED = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "selfpaced_*.npz")))
IIC = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "iic_*.npz")) if "survey" not in p)
IID = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "iid_*.npz")))


def _kw(g):
    kw = {}
    if "mask" in g.files:
        kw["mask"] = g["mask"]
    elif "labels" in g.files:
        kw["target"] = g["labels"].tolist()
    return kw


def test_survey_known_answers():
    g = load_golden("supcon_survey")
    lab = g["labels"].tolist()
    assert C.supcon(g["f1"], g["f2"], target=lab)["loss"] == pytest.approx(5.235924850353655, rel=1e-12)
    assert C.supcon(g["f1"], g["f2"])["loss"] == pytest.approx(float(g["simclr_survey"]), rel=1e-12)
    assert C.supcon(g["f1"], g["f2"], target=list(range(64)))["loss"] == pytest.approx(float(g["simclr_survey"]), rel=1e-12)
    assert C.supcon(g["f1"], g["f2"], target=lab, exclude_other_pos=True)["loss"] == pytest.approx(
        float(g["exclude_survey"]), rel=1e-8)
    g = load_golden("iic_survey")
    assert M.iid_segmentation_loss(g["x"], g["y"], padding=1)["loss"] == pytest.approx(-0.2478573718202518, rel=1e-11)
    assert M.iid_segmentation_loss(g["x"], g["y"], padding=0)["loss"] == pytest.approx(float(g["loss_pad0"]), rel=1e-10)


@pytest.mark.parametrize("name", SUPCON)
def test_supcon_golden(name):
    g = load_golden(name)
    o = C.supcon(g["f1"], g["f2"], t=float(g["temperature"]), exclude_other_pos=bool(g["exclude"]), **_kw(g))
    rel = 1e-7 if bool(g["exclude"]) else 1e-11     # the reference's .float() cast of the counts (contrastive.py:88)
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=rel)
    np.testing.assert_array_equal(o["pos_mask"].astype(np.uint8), g["pos_mask"])     # bit-exact
    np.testing.assert_array_equal(o["neg_mask"].astype(np.uint8), g["neg_mask"])
    scale = np.abs(g["grad_f1"]).max()
    np.testing.assert_allclose(o["grad_f1"], g["grad_f1"], rtol=0, atol=scale * (1e-6 if bool(g["exclude"]) else 1e-10))
    np.testing.assert_allclose(o["grad_f2"], g["grad_f2"], rtol=0, atol=scale * (1e-6 if bool(g["exclude"]) else 1e-10))
    if "sim_exp" in g.files:
        np.testing.assert_allclose(o["sim_exp"], g["sim_exp"], rtol=1e-11)
        np.testing.assert_allclose(o["sim_logits"], g["sim_logits"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", SELFPACED)
def test_selfpaced_golden(name):
    g = load_golden(name)
    o = C.selfpaced_supcon(g["f1"], g["f2"], t=float(g["temperature"]), weight_update=str(g["mode"]),
                           gamma=float(g["gamma"]), correct_grad=bool(g["correct_grad"]), **_kw(g))
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=1e-11)
    assert o["downgrade_ratio"] == pytest.approx(float(g["downgrade_ratio"]), rel=1e-12)
    np.testing.assert_allclose(o["sp_mask"], g["sp_mask"], rtol=1e-10, atol=1e-12)
    scale = np.abs(g["grad_f1"]).max()
    np.testing.assert_allclose(o["grad_f1"], g["grad_f1"], rtol=0, atol=scale * 1e-10)
    np.testing.assert_allclose(o["grad_f2"], g["grad_f2"], rtol=0, atol=scale * 1e-10)


def test_selfpaced_large_gamma_equals_supcon():
    # the reference's own self-check (contrastive.py:228-248): soft weights with a huge gamma == SupConLoss1
    g = load_golden("supcon_labels5")
    lab = g["labels"].tolist()
    a = C.selfpaced_supcon(g["f1"], g["f2"], target=lab, weight_update="soft", gamma=1e10)
    b = C.supcon(g["f1"], g["f2"], target=lab)
    assert a["loss"] == pytest.approx(b["loss"], rel=1e-8)


@pytest.mark.parametrize("name", IIC)
def test_iic_golden(name):
    g = load_golden(name)
    o = M.iid_segmentation_loss(g["x"], g["y"], lamda=float(g["lamda"]), padding=int(g["padding"]), eps=float(g["eps"]),
                                symmetric=bool(g["symmetric"]), mask=g["mask"] if "mask" in g.files else None)
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=1e-10, abs=1e-14)
    np.testing.assert_allclose(o["joint"], g["joint"], rtol=1e-10)
    for k in ("grad_x", "grad_y"):
        np.testing.assert_allclose(o[k], g[k], rtol=0, atol=np.abs(g[k]).max() * 1e-8)


IIC_LOGITS = ["logits_iic_pad1_T05", "logits_iic_pad1_sym_T2", "logits_iic_pad0_T1"]


@pytest.mark.parametrize("name", IIC_LOGITS)
def test_iic_from_logits_golden(name):
    """the reference's own SoftmaxWithT tail + sub-head mean of IIDSegmentationLoss, gradients w.r.t. the logits
    (tests/golden/make_golden.py logits_case): pins oracle.discrete_mi_np.iid_segmentation_loss_from_logits"""
    g = load_golden(name)
    o = M.iid_segmentation_loss_from_logits(list(g["logits_x"]), list(g["logits_y"]), float(g["T"]), padding=int(g["padding"]),
                                            symmetric=bool(g["symmetric"]))
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=1e-10, abs=1e-14)
    for k in ("grad_x", "grad_y"):
        np.testing.assert_allclose(o[k], g[k], rtol=0, atol=np.abs(g[k]).max() * 1e-8)


def test_iic_negative_padding_raises():
    g = load_golden("iic_pad1")
    with pytest.raises(ValueError):
        M.iid_segmentation_loss(g["x"], g["y"], padding=-1)


@pytest.mark.parametrize("name", IID)
def test_iid_golden(name):
    g = load_golden(name)
    o = M.iid_loss(g["x"], g["y"], lamb=float(g["lamb"]))
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=1e-11)
    assert o["loss_no_lamb"] == pytest.approx(float(g["loss_no_lamb"]), rel=1e-11)
    np.testing.assert_allclose(o["p_i_j"], g["p_i_j"], rtol=1e-12)
    np.testing.assert_allclose(o["grad_x"], g["grad_x"], rtol=0, atol=np.abs(g["grad_x"]).max() * 1e-10)
    np.testing.assert_allclose(o["grad_y"], g["grad_y"], rtol=0, atol=np.abs(g["grad_y"]).max() * 1e-10)


def test_label_generators_golden():
    cases = json.load(open(os.path.join(GOLDEN, "labels.json")))
    assert Lb.label_encode(cases["partition"]["in"]) == cases["partition"]["out"]
    assert Lb.label_encode(cases["patient"]["in"]) == cases["patient"]["out"]
    assert Lb.cycle_labels(cases["cycle"]["in"]) == cases["cycle"]["out"]
    assert Lb.self_labels(cases["self"]["in"]) == cases["self"]["out"]
    groups = [f"{p}_{e}" for p, e in zip(cases["patient"]["in"], cases["cycle"]["in"])]
    assert Lb.get_label("patient", "acdc", cases["partition"]["in"], groups) == cases["patient"]["out"]
    assert Lb.get_label("cycle", "acdc_lv", cases["partition"]["in"], groups) == cases["cycle"]["out"]
    assert Lb.get_label("partition", "prostate", cases["partition"]["in"], groups) == cases["partition"]["out"]


def test_region_coordinates_golden():
    for c in json.load(open(os.path.join(GOLDEN, "regions.json"))):
        got = Lb.region_coordinates(c["b"], c["h"], c["w"], c["point_nums"], c["seed"])
        assert got.tolist() == c["coords"]


# ---------------- the chunked / C forms against the literal numpy oracle ----------------
def test_chunked_numpy_matches_literal():
    g = load_golden("supcon_partition3")
    lab = g["labels"]
    lit = C.supcon(g["f1"], g["f2"], target=lab.tolist())
    z = np.concatenate([g["f1"], g["f2"]])
    ch = C.supcon_chunked(z, np.tile(lab, 2), chunk=37)
    assert ch["loss"] == pytest.approx(lit["loss"], rel=1e-12)
    np.testing.assert_allclose(ch["grad"], np.concatenate([lit["grad_f1"], lit["grad_f2"]]), rtol=0,
                               atol=np.abs(lit["grad_f1"]).max() * 1e-10)


@pytest.mark.parametrize("name", ["supcon_partition3", "supcon_simclr", "supcon_d256", "supcon_tiny", "supcon_biglabels"])
@pytest.mark.parametrize("prec", [0, 1])
def test_c_supcon_matches_golden(name, prec):
    g = load_golden(name)
    n = g["f1"].shape[0]
    if "labels" in g.files:
        # the reference compares the float32 images of the labels (contrastive.py:40): rank them the same way
        lab32 = np.asarray(g["labels"].tolist(), dtype=np.float32)
        lab = np.unique(lab32, return_inverse=True)[1].astype(np.int32)
    else:
        lab = np.arange(n, dtype=np.int32)
    z = np.concatenate([g["f1"], g["f2"]]).astype(np.float32)
    o = c_oracle.supcon_fwd_bwd(z, np.tile(lab, 2), t=float(g["temperature"]), prec=prec)
    # inputs are rounded to float32 on the way in: 1e-6 relative is the float32 input-rounding floor
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=2e-6)
    ref = np.concatenate([g["grad_f1"], g["grad_f2"]])
    np.testing.assert_allclose(o["grad"], ref, rtol=0, atol=np.abs(ref).max() * (2e-5 if prec == 0 else 5e-6))


@pytest.mark.parametrize("name", IIC)
@pytest.mark.parametrize("prec", [0, 1])
def test_c_iic_matches_golden(name, prec):
    g = load_golden(name)
    if "mask" in g.files:
        pytest.skip("the C port takes pre-masked maps")
    o = c_oracle.iic_fwd_bwd(g["x"], g["y"], int(g["padding"]), symmetric=bool(g["symmetric"]), lamda=float(g["lamda"]),
                             eps=float(g["eps"]), prec=prec)
    assert o["loss"] == pytest.approx(float(g["loss"]), rel=2e-5, abs=1e-9)
    np.testing.assert_allclose(o["joint"], g["joint"], rtol=1e-5)
    for k in ("grad_x", "grad_y"):
        np.testing.assert_allclose(o[k], g[k], rtol=0, atol=np.abs(g[k]).max() * 2e-5)
