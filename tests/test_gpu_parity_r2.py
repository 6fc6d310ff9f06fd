"""GPU parity tests added in round 2 ("Parity tests for what is benchmarked"):

* the BENCHMARKED NAP -- D = 1728, all layers (novelty_detection.py:56-57,69-70 -> utils/metric.py:183-238), eigen and
  triangular factor, every tensor-core mode -- against the reference's golden scores and the fp64 value of its formula;
* single-layer (well-conditioned) NAP within 1e-4 of the fp64 closed form;
* the pipelined bulk path of mmad_score_host (>= 2048 rows, several host chunks) against the oracle;
* the VIB decorator kernel (decorators/variational_info_bottleneck.py:19-42) against the reference golden: k > 1
  expansion, deterministic branch, ValueError / NotImplementedError;
* trained autoencoders at D in {64, 128, 1728} over 65 536 windows: every window within 1e-4 in the benchmarked mode;
* state invalidation: new weights / a precision switch drop cached launch graphs and the installed NAP fit.
"""
import argparse
import functools

import numpy as np
import pytest
import torch

from conftest import load_golden
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

pytestmark = pytest.mark.gpu


def _model(D, btl, nl, sd, precision):
    from icra2021_multimodal_ad_b200.model_builder import get_model
    m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=precision))
    m.load_state_dict(sd)
    return m.eval()


# ------------------------------------------------------------------------------------------------------------
# NAP, D = 1728, all layers: what bench.py times
# ------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _nap_case(name, sel):
    """(golden, state dict, train / test windows, labels, fp64 NAP of the oracle's diffs) -- computed once."""
    from oracle import rapp_oracle as RO
    g = load_golden(name)
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    sd = synth_state_dict(D, btl, nl, seed)
    xtr, _ = synth_windows(g["n_tr"], D, seed + 1, anomaly_rate=0.0)
    xte, yte = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
    lo, hi = sel
    dtr, dte = RO.get_diffs(xtr, sd), RO.get_diffs(xte, sd)
    truth = RO.nap_score_fp64(RO.concat_diffs(dtr, lo, hi), RO.concat_diffs(dte, lo, hi))
    return g, sd, xtr, xte, yte.numpy().astype(bool), truth


@functools.lru_cache(maxsize=None)
def _nap_full_case():
    from oracle import rapp_oracle as RO
    g = load_golden("nap_D1728_full.pt")
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    sd = synth_state_dict(D, btl, nl, seed)
    xtr, _ = synth_windows(g["n_tr"], D, seed + 1, anomaly_rate=0.0)
    xte, yte = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
    truth = RO.nap_score_fp64(RO.concat_diffs(RO.get_diffs(xtr, sd)), RO.concat_diffs(RO.get_diffs(xte, sd)))
    return g, sd, xtr, xte, yte.numpy().astype(bool), truth


@pytest.mark.parametrize("factor", ["hybrid", "eigen", "triangular"])
@pytest.mark.parametrize("precision", ["fp32", "f16x3", "f16f8"])
def test_nap_all_layers_protocol_headline_width(precision, factor):
    """SURVEY 8c-4 at the BENCHMARKED shape: D = 1728, all six diffs (D' = 5482), fit set of N_tr = 6144 >= D' rows so
    that K = D' like bench.py's fit, golden scores from the unmodified reference (tests/golden/make_golden.py nap_full).
    The selection is rank deficient by construction (F5: d_5 = W_5 d_4; measured: ~770 of the 5482 singular values sit
    at the fp32 noise floor, 1e-8 of the largest), so the reference's own fp32 result is 51 % away (median) from the fp64
    value of its formula while ranking the windows like it (Spearman 0.974, AUROC 0.760 vs 0.744).
    Required of the shipped factorisations ("hybrid" = default, "eigen"): median error against the fp64 value no worse
    than the reference's in fp32 mode and within 1.25x of it in the tensor-core modes (their diffs carry ~1e-6 instead
    of ~1e-7 relative rounding noise, which the weak directions amplify), rank agreement within 0.02 of the reference's,
    AUROC within 0.02 of the reference's.  The pure triangular factor spreads the weak directions' noise over every
    output (engine.nap_fit_from_stats) and is only held to the AUROC and a loose rank bar here -- it is the right
    choice for well-conditioned selections (tests below), not for this one."""
    from scipy.stats import spearmanr
    from icra2021_multimodal_ad_b200.utils import metric as M
    g, sd, xtr, xte, y, truth = _nap_full_case()
    D, btl, nl = g["D"], g["btl"], g["n_layers"]
    ref = g["nap"]["score"].numpy().astype(np.float64)
    eng = _model(D, btl, nl, sd, precision).engine()
    eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False, factor=factor)
    new = eng.score(xte.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
    ok = np.isfinite(new) & np.isfinite(ref)
    assert ok.mean() > 0.99
    err_new = np.median(np.abs(new[ok] - truth[ok]) / truth[ok])
    err_ref = np.median(np.abs(ref[ok] - truth[ok]) / truth[ok])
    rho_new = spearmanr(new[ok], truth[ok]).correlation
    rho_ref = spearmanr(ref[ok], truth[ok]).correlation
    auc_new = M.get_auc_roc(new.astype(np.float32), y)
    auc_ref, auc_truth = g["nap"]["metrics"][0], M.get_auc_roc(truth.astype(np.float32), y)
    print("NAP all layers D=1728 N_tr=6144 [%s %s]: err_new %.3g err_ref %.3g rho_new %.4f rho_ref %.4f auroc new %.4f ref %.4f fp64 %.4f"
          % (precision, factor, err_new, err_ref, rho_new, rho_ref, auc_new, auc_ref, auc_truth))
    if factor == "triangular":
        assert rho_new >= 0.8 and abs(auc_new - auc_ref) <= 0.03
        return
    assert err_new <= err_ref * (1.05 if precision == "fp32" else 1.25)
    assert rho_new >= rho_ref - 0.02
    assert abs(auc_new - auc_ref) <= 0.02


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "f16f8"])
def test_nap_all_layers_rank_limited_fit(precision):
    """The other committed reference golden, score_D1728.pt: N_tr = 2048 < D' = 5482.  The centred fit matrix has rank
    N_tr - 1, its last component is a pure rounding-noise direction with variance ~0, and the score is dominated by it:
    the REFERENCE's own scores are uncorrelated with the fp64 value of its formula (measured rho_ref = -0.0015,
    err_ref = 1.0) and its AUROC is at chance level (0.53), so no per-window comparison means anything here.  What can
    be held: finite scores, and an AUROC at chance level like the reference's (measured 0.43 .. 0.55 over the modes)."""
    from icra2021_multimodal_ad_b200.utils import metric as M
    g, sd, xtr, xte, y, _ = _nap_case("score_D1728.pt", (0, 1))        # cheap selection: only the inputs are used here
    D, btl, nl = g["D"], g["btl"], g["n_layers"]
    eng = _model(D, btl, nl, sd, precision).engine()
    eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False)
    new = eng.score(xte.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy()
    assert np.isfinite(new).mean() > 0.99
    auc_new, auc_ref = M.get_auc_roc(new, y), g["nap"]["0:7"]["metrics"][0]
    print("NAP all layers D=1728 N_tr=2048 [%s]: auroc new %.4f ref %.4f" % (precision, auc_new, auc_ref))
    assert abs(auc_new - 0.5) <= 0.12 and abs(auc_ref - 0.5) <= 0.12      # both at chance level: the score is noise


@pytest.mark.parametrize("factor", ["eigen", "triangular"])
@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("name,sels", [("score_D64.pt", ((0, 1), (1, 2))), ("score_D1728.pt", ((0, 1), (5, 6)))])
def test_nap_single_layers_within_1e4_of_fp64(name, sels, precision, factor):
    """NAP on well-conditioned selections within 1e-4 relative per window.  The yardstick is the fp64
    closed form on the oracle's diffs: the reference's own fp32 SVD carries ~1e-4 of noise (its golden scores are held
    to 1e-3 in test_nap_single_layers_match_reference), the fp64 value does not.  Selections: d_0 and d_1 alone at
    D = 64, d_0 and d_5 alone at D = 1728 (SURVEY: cond ~ 9 / 97 / 71).  NOT d_5 at D = 64: that network widens
    (64 -> ... -> 92 -> 100), so d_5 = W_5 d_4 has rank <= 92 of 100 -- rank deficient like the all-layers case."""
    for sel in sels:
        g, sd, xtr, xte, y, truth = _nap_case(name, sel)
        D, btl, nl = g["D"], g["btl"], g["n_layers"]
        eng = _model(D, btl, nl, sd, precision).engine()
        eng.nap_fit(xtr.cuda(), sel[0], sel[1], distributed=False, factor=factor)
        s = eng.score(xte.cuda(), sel[0], sel[1], base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
        err = np.abs(s - truth) / truth
        print(f"NAP {name} {sel} [{precision} {factor}]: max {err.max():.2e} median {np.median(err):.2e}")
        assert err.max() < 1e-4, (sel, precision, factor)


def test_nap_rotation_with_fp8_cross_terms_option():
    """mmad_set_option("nap_passes", 4): the NAP rotation of the f16x3 mode as one fp16 hi*hi MMA + one fp8 MMA carrying both
    cross terms (2 tensor-work units per product instead of 3).  Opt-in, not the benchmarked default.  Held to the SAME bars
    as the default: 1e-4 per window on well-conditioned selections, the all-layers protocol at the benchmarked width.
    Changing the option drops the installed fit (its variances carry the old arithmetic's noise)."""
    from scipy.stats import spearmanr
    from icra2021_multimodal_ad_b200.utils import metric as M
    from icra2021_multimodal_ad_b200._lib import MmadError
    for name, sel in (("score_D1728.pt", (0, 1)), ("score_D1728.pt", (5, 6)), ("score_D64.pt", (0, 1))):
        g, sd, xtr, xte, y, truth = _nap_case(name, sel)
        eng = _model(g["D"], g["btl"], g["n_layers"], sd, "f16x3").engine()
        eng.nap_fit(xtr.cuda(), sel[0], sel[1], distributed=False)
        eng.set_option("nap_passes", 4)
        with pytest.raises(MmadError):
            eng.score(xte.cuda(), sel[0], sel[1], base=False, sap=False, nap=True)
        eng.nap_fit(xtr.cuda(), sel[0], sel[1], distributed=False)
        s = eng.score(xte.cuda(), sel[0], sel[1], base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
        err = np.abs(s - truth) / truth
        print(f"NAP fp8 cross terms {name} {sel}: max {err.max():.2e} median {np.median(err):.2e}")
        assert err.max() < 1e-4, (name, sel)
    g, sd, xtr, xte, y, truth = _nap_full_case()
    D, btl, nl = g["D"], g["btl"], g["n_layers"]
    ref = g["nap"]["score"].numpy().astype(np.float64)
    eng = _model(D, btl, nl, sd, "f16x3").engine()
    eng.set_option("nap_passes", 4)
    eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False)
    new = eng.score(xte.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
    ok = np.isfinite(new) & np.isfinite(ref)
    err_new = np.median(np.abs(new[ok] - truth[ok]) / truth[ok])
    err_ref = np.median(np.abs(ref[ok] - truth[ok]) / truth[ok])
    rho_new, rho_ref = spearmanr(new[ok], truth[ok]).correlation, spearmanr(ref[ok], truth[ok]).correlation
    auc_new, auc_ref = M.get_auc_roc(new.astype(np.float32), y), g["nap"]["metrics"][0]
    print("NAP fp8 cross terms, all layers D=1728: err %.3g (ref %.3g) rho %.4f (ref %.4f) auroc %.4f (ref %.4f)"
          % (err_new, err_ref, rho_new, rho_ref, auc_new, auc_ref))
    assert err_new <= err_ref * 1.25 and rho_new >= rho_ref - 0.02 and abs(auc_new - auc_ref) <= 0.02


# ------------------------------------------------------------------------------------------------------------
# mmad_score_host, bulk (pipelined) path
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "f16x3", "f16f8"])
@pytest.mark.parametrize("D,btl,nl,n", [(300, 17, 2, 2 * 18944 + 1301), (1728, 100, 5, 18944 + 2100)])
def test_score_host_pipelined_path_matches_oracle(D, btl, nl, n, precision):
    """> 2048 rows: double-buffered H2D / compute / D2H over several host chunks with a ragged last chunk; pinned and
    pageable input give the same scores; base, SAP and NAP against the oracle / the device-resident call."""
    from oracle import rapp_oracle as RO
    if precision == "fp32" and D == 1728:
        n = 2048 + 700          # CUDA-core mode: keep the test short, still the bulk path
    sd = synth_state_dict(D, btl, nl, 77)
    x, _ = synth_windows(n, D, 5)
    ref = RO.get_diffs(x, sd)
    eng = _model(D, btl, nl, sd, precision).engine()
    xtr, _ = synth_windows(max(2 * D, 1024), D, 6, anomaly_rate=0.0)
    eng.nap_fit(xtr.cuda(), 0, 1, distributed=False)
    tol = 1e-4 if (precision != "f16f8" or D > 128) else 5e-4
    xp = x.pin_memory()
    o = eng.score_host(xp.numpy(), 0, nl + 1, nap=False)
    np.testing.assert_allclose(o["sap"], RO.sap_score(ref), rtol=tol)
    np.testing.assert_allclose(o["base"], RO.recon_score(ref[0]), rtol=tol)
    o2 = eng.score_host(x.numpy().copy(), 0, nl + 1, nap=False)          # pageable
    assert np.array_equal(o2["sap"], o["sap"]) and np.array_equal(o2["base"], o["base"])
    # NAP through the host path == NAP through the device path (same kernels, same chunk arithmetic)
    oh = eng.score_host(xp.numpy(), 0, 1, base=False, sap=False, nap=True)["nap"]
    od = eng.score(x.cuda(), 0, 1, base=False, sap=False, nap=True)["nap"].cpu().numpy()
    np.testing.assert_allclose(oh, od, rtol=2e-5)
    # strict mode rejects pageable bulk input
    from icra2021_multimodal_ad_b200._lib import MmadError
    eng.set_option("require_pinned", 1)
    with pytest.raises(MmadError):
        eng.score_host(x.numpy().copy(), 0, nl + 1)
    eng.score_host(xp.numpy(), 0, nl + 1)
    eng.set_option("require_pinned", 0)


# ------------------------------------------------------------------------------------------------------------
# VIB decorator kernel
# ------------------------------------------------------------------------------------------------------------
def test_vib_decorator_kernel_matches_reference_golden():
    from icra2021_multimodal_ad_b200.modules.fc_module import FCModule
    g = load_golden("vib_D64.pt")
    k = g["k"]
    enc = FCModule(input_size=64, output_size=40, hidden_sizes=[56, 48], use_batch_norm=True, act="leakyrelu", last_act=None)
    enc.load_state_dict(g["sd"])
    enc = enc.cuda().eval()
    x = g["x"].cuda()
    rel = lambda a, b: float((a.cpu() - b).abs().max() / b.abs().max())  # noqa: E731
    with torch.no_grad():
        assert rel(enc(x), g["plain"]) < 1e-5
        r = enc(x, distribution="normal", k=k, eps=g["eps"].cuda())
        assert tuple(r["z"].shape) == tuple(g["z"].shape) and tuple(r["mu"].shape) == tuple(g["mu"].shape)
        assert rel(r["z"], g["z"]) < 1e-5 and rel(r["mu"], g["mu"]) < 1e-5 and rel(r["logvar"], g["logvar"]) < 1e-5
        det = enc(x, distribution="normal", k=k, stochastic_inference=False)
        assert tuple(det["z"].shape) == tuple(g["z_det"].shape) and rel(det["z"], g["z_det"]) < 1e-5
        for j in range(k):                                     # deterministic branch: mu repeated k times
            assert torch.equal(det["z"][j], det["mu"])
        # stochastic_inference=True without eps draws device noise: right shape, mean mu, not mu itself
        s = enc(x, distribution="normal", k=k)
        assert tuple(s["z"].shape) == tuple(g["z"].shape) and not torch.equal(s["z"][0], s["mu"])
        with pytest.raises(ValueError):
            enc(x, distribution="normal", k=0)
        with pytest.raises(NotImplementedError):
            enc(x, distribution="laplace")
        # the (k, B, h) code goes through a decoder FCModule like the reference's (BatchNorm flattens, fc_layer.py:41-43)
        dec = FCModule(input_size=20, output_size=64, hidden_sizes=[48], use_batch_norm=True).cuda().eval()
        out = dec(r["z"])
        assert tuple(out.shape) == (k, x.shape[0], 64)
        flat = dec(r["z"].reshape(-1, 20)).reshape(k, x.shape[0], 64)
        assert torch.equal(out, flat)


# ------------------------------------------------------------------------------------------------------------
# trained models, 65 536 windows
# ------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _trained_case(D, steps, n):
    from oracle import rapp_oracle as RO
    btl, nl = 100, 5
    sd = synth_state_dict(D, btl, nl, 0)
    xtr, _ = synth_windows(256 * 8, D, 7, anomaly_rate=0.0)
    opt = {}
    for i in range(steps):
        RO.train_step(xtr[(i % 8) * 256:(i % 8 + 1) * 256], sd, opt)
    x, _ = synth_windows(n, D, 1236)
    ref = RO.get_diffs(x, sd, batch_size=4096)
    return sd, x, RO.sap_score(ref).astype(np.float64), RO.recon_score(ref[0]).astype(np.float64)


# (median, max) relative error per window; the benchmarked mode (f16x3) must hold max <= 1e-4 at every width
TRAINED_BARS = {"fp32": (2e-6, 2e-5), "f16x3": (2e-5, 1e-4)}


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("D,steps", [(64, 100), (128, 100), (1728, 60)])
def test_trained_models_65536_windows(D, steps, precision):
    """a TRAINED autoencoder (diffs are small differences of large activations) over 65 536 windows,
    every window's base and SAP score within 1e-4 of the oracle's fp32 scores in the mode bench.py reports."""
    n = 65536 if precision != "fp32" or D < 1728 else 8192          # CUDA-core mode at D = 1728: 8192 windows
    sd, x, sap_o, base_o = _trained_case(D, steps, 65536)
    m = _model(D, 100, 5, sd, precision)
    o = m.engine().score(x[:n].cuda(), 0, 6)
    b_med, b_max = TRAINED_BARS[precision]
    for name, got, want in (("sap", o["sap"], sap_o[:n]), ("base", o["base"], base_o[:n])):
        err = np.abs(got.cpu().numpy() - want) / want
        print(f"trained D={D} n={n} {precision} {name}: max {err.max():.2e} p99.9 {np.quantile(err, 0.999):.2e} median {np.median(err):.2e}")
        assert np.median(err) < b_med and err.max() < b_max, (precision, name)


# ------------------------------------------------------------------------------------------------------------
# state invalidation
# ------------------------------------------------------------------------------------------------------------
def test_new_weights_drop_cached_graphs_and_nap_fit():
    """f16f8: the power-of-two fp8 weight scale is a kernel argument baked into the cached latency-path graph; weights
    scaled by 4 cross a power of two.  After set_layer the same host call must score the NEW model."""
    from icra2021_multimodal_ad_b200._lib import MmadError
    D, btl, nl = 300, 17, 2
    sd = synth_state_dict(D, btl, nl, 3)
    x, _ = synth_windows(256, D, 11)
    m8 = _model(D, btl, nl, sd, "f16f8")
    m32 = _model(D, btl, nl, sd, "fp32")
    a = m8.engine().score_host(x.numpy(), 0, nl + 1)
    np.testing.assert_allclose(a["sap"], m32.engine().score_host(x.numpy(), 0, nl + 1)["sap"], rtol=5e-4)
    xtr, _ = synth_windows(1024, D, 6, anomaly_rate=0.0)
    eng = m8.engine()
    eng.nap_fit(xtr.cuda(), 0, 1, distributed=False)
    eng.score(x.cuda(), 0, 1, nap=True)
    sd4 = {k: (v * 4 if k.endswith("layer.weight") else v) for k, v in sd.items()}
    m8.load_state_dict(sd4)
    m32.load_state_dict(sd4)
    b = m8.engine().score_host(x.numpy(), 0, nl + 1)
    np.testing.assert_allclose(b["sap"], m32.engine().score_host(x.numpy(), 0, nl + 1)["sap"], rtol=5e-4)
    assert not np.allclose(a["sap"], b["sap"], rtol=1e-2)
    # the NAP fit belonged to the old weights
    assert m8.engine().nap_range is None
    with pytest.raises(MmadError):
        m8.engine().score(x.cuda(), 0, 1, nap=True)
    # ... and to the old arithmetic
    eng = m8.engine()
    eng.nap_fit(xtr.cuda(), 0, 1, distributed=False)
    eng.set_precision("f16x3")
    assert eng.nap_range is None
    with pytest.raises(MmadError):
        eng.score(x.cuda(), 0, 1, nap=True)
    with pytest.raises(MmadError):
        eng.set_option("no_such_option", 1)


# ------------------------------------------------------------------------------------------------------------
# realtime entry point
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,btl,nl", [(1728, 100, 5), (128, 100, 5), (300, 17, 2)])
def test_realtime_detecter_test_call(D, btl, nl, tmp_path):
    """test_file/realtime_tester.py:291-304: ``detecter.test(model, fusion_representation, config, nap=False)`` -> one SAP
    score per window of the (batch_size = 10, D) matrix; host matrices run in one launch (stream kernel), every batch
    size 1..64, deterministic; nap=True scores from a NAP-fit checkpoint."""
    from icra2021_multimodal_ad_b200.NoveltyDetecter import NoveltyDetecter
    from icra2021_multimodal_ad_b200.novelty_detection import NoveltyDetecter as Trainer
    from icra2021_multimodal_ad_b200._lib import lib
    from oracle import rapp_oracle as RO
    sd = synth_state_dict(D, btl, nl, 9)
    cfg = argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision="f16x3", batch_size=10,
                             start_layer_index=0, end_layer_index=-1)
    m = _model(D, btl, nl, sd, "f16x3")
    det = NoveltyDetecter(cfg)
    for n in (1, 2, 4, 5, 8, 9, 10, 12, 13, 16, 17, 40, 64):
        x, _ = synth_windows(n, D, 200 + n)
        want = RO.sap_score(RO.get_diffs(x, sd))
        det.test(m, x, cfg, nap=False)                       # first call packs the weights / builds the plan
        l0 = lib().mmad_launch_count()
        got = det.test(m, x, cfg, nap=False)
        if n <= 16:
            assert lib().mmad_launch_count() - l0 == 1      # one kernel launch per realtime call (<= 16 windows)
        assert isinstance(got, list) and len(got) == n
        np.testing.assert_allclose(np.asarray(got), want, rtol=5e-5 if n <= 16 else 1e-4)
        assert det.test(m, x.numpy(), cfg, nap=False) == got
        np.testing.assert_allclose(np.asarray(det.test(m, x.cuda(), cfg, nap=False)), want, rtol=1e-4)
    buf = det.window_buffer(m, cfg)
    assert buf.shape == (64, D) and buf.dtype == np.float32
    x, _ = synth_windows(10, D, 77)
    buf[:10] = x.numpy()
    np.testing.assert_allclose(np.asarray(det.test(m, buf[:10], cfg)), RO.sap_score(RO.get_diffs(x, sd)), rtol=5e-5)
    # NAP from a checkpoint (well-conditioned selection d_0)
    cfg.end_layer_index = nl                    # diffs[0:1]
    cfg.nap_fit = str(tmp_path / "nap_fit.pt")
    xtr, _ = synth_windows(max(2 * D, 600), D, 6, anomaly_rate=0.0)
    xva, _ = synth_windows(64, D, 7, anomaly_rate=0.0)
    xte, yte = synth_windows(100, D, 8)
    fast = Trainer(cfg).score_fast(m, xtr, xva, xte, yte.numpy().astype(int))
    m.engine().load_state_dict(m.state_dict())          # drop the fit; the realtime detector reloads it from the checkpoint
    got = det.test(m, xte[:10], cfg, nap=True)
    np.testing.assert_allclose(np.asarray(got), fast["nap"]["score"][:10].cpu().numpy(), rtol=1e-3)


# ------------------------------------------------------------------------------------------------------------
# small-net fused chain (north_star item 1; VERDICT r1 row "small-net fused chain")
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("D,btl,nl", [(64, 100, 5), (128, 100, 5), (93, 10, 3), (48, 100, 5), (128, 128, 2)])
def test_smallnet_fused_chain_matches_oracle(D, btl, nl, precision):
    """Per-modality models (utils/data_loaders.py:16-29: force_torque 64, mic 128; every width <= 128): base and SAP come
    from ONE fused kernel per call -- CUDA cores in the fp32 mode (smallnet.cu), tensor cores with weights and activations
    in shared memory and accumulators in TMEM in the f16x3 mode (smallnet_tc.cu, calls of more than 64 rows); rows must be
    16-byte aligned (D % 4 == 0), other cases keep the per-layer kernels.  Against the oracle, against the per-layer kernels (mmad_set_option
    smallnet = 0), every layer selection, ragged row counts, odd widths."""
    from icra2021_multimodal_ad_b200._lib import lib
    from oracle import rapp_oracle as RO
    sd = synth_state_dict(D, btl, nl, 13)
    eng = _model(D, btl, nl, sd, precision).engine()
    tol = 2e-5 if precision == "fp32" else 1e-4
    for n in (1, 31, 32, 33, 127, 128, 129, 257, 20011):
        # fp32: the fused CUDA-core kernel at every row count; f16x3: the fused tensor-core kernel above 64 rows
        fused = D % 4 == 0 and (precision == "fp32" or n > 64)
        x, _ = synth_windows(n, D, 300 + n)
        ref = RO.get_diffs(x, sd)
        xd = x.cuda()
        eng.score(xd, 0, nl + 1)
        l0 = lib().mmad_launch_count()
        o = eng.score(xd, 0, nl + 1)
        assert (lib().mmad_launch_count() - l0 == 1) == fused, (n, precision)
        np.testing.assert_allclose(o["sap"].cpu().numpy(), RO.sap_score(ref), rtol=tol)
        np.testing.assert_allclose(o["base"].cpu().numpy(), RO.recon_score(ref[0]), rtol=tol)
        if n == 257:
            for lo, hi in ((0, 1), (1, 2), (1, nl), (nl, nl + 1), (1, nl + 1)):
                s = eng.score(xd, lo, hi)
                np.testing.assert_allclose(s["sap"].cpu().numpy(), RO.sap_score(ref, lo, hi), rtol=tol)
                np.testing.assert_allclose(s["base"].cpu().numpy(), RO.recon_score(ref[0]), rtol=tol)
            eng.set_option("smallnet", 0)
            l0 = lib().mmad_launch_count()
            p = eng.score(xd, 0, nl + 1)
            assert lib().mmad_launch_count() - l0 > 1                      # the per-layer kernels of the handle's mode
            eng.set_option("smallnet", 1)
            np.testing.assert_allclose(o["sap"].cpu().numpy(), p["sap"].cpu().numpy(), rtol=1e-5 if precision == "fp32" else 1e-4)
    # rows that are not 16-byte aligned / strided views fall back to the per-layer path, same scores
    x, _ = synth_windows(100, D, 5)
    big = torch.zeros(100, D + 3).cuda()
    big[:, :D] = x.cuda()
    a = eng.score(x.cuda(), 0, nl + 1)["sap"].cpu().numpy()
    b = eng.score(big[:, :D], 0, nl + 1)["sap"].cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-5 if precision == "fp32" else 1e-4)
    # host entry points: bulk (pipelined chunks) and graph-replay sizes
    for n in (40, 3000, 40000):
        x, _ = synth_windows(n, D, 7 + n)
        h = eng.score_host(x.numpy(), 0, nl + 1)
        np.testing.assert_allclose(h["sap"], RO.sap_score(RO.get_diffs(x, sd)), rtol=tol)
