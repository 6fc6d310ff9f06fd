"""CPU check of the operand-split arithmetic behind the tensor-core modes (scripts/emulate_split_f8.py): with exact
products and accumulation, F16X3 (hi/lo fp16 pairs) is fp32-equivalent, F16F8 (fp16 hi*hi + e4m3 cross terms, per-tensor
power-of-two weight scale) stays within its stated bound, and both beat the one- and two-pass fp16 variants.  What the
hardware adds on top (truncating accumulation) is measured on the GPU (tests/test_gpu_f16f8.py)."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _emu():
    spec = importlib.util.spec_from_file_location("emulate_split_f8", os.path.join(ROOT, "scripts", "emulate_split_f8.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_split_modes_on_a_small_trained_model():
    from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
    from oracle import rapp_oracle as RO
    E = _emu()
    D = 64
    sd = synth_state_dict(D, 100, 5, 0)
    xtr, _ = synth_windows(256 * 4, D, 7, anomaly_rate=0.0)
    opt = {}
    for i in range(40):
        RO.train_step(xtr[(i % 4) * 256:(i % 4 + 1) * 256], sd, opt)
    x, _ = synth_windows(512, D, 1236)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    ref = torch.cat(E.diffs(x.double(), sd64, "fp32"), 1).pow(2).mean(1)
    err = {}
    for mode in ("f16x3", "f16f8", "f16x2", "f16"):
        sap = torch.cat(E.diffs(x, sd, mode), 1).double().pow(2).mean(1)
        err[mode] = float(((sap - ref).abs() / ref).max())
    assert err["f16x3"] < 5e-6, err
    assert err["f16f8"] < 5e-4, err
    assert err["f16f8"] < err["f16x2"] < err["f16"], err


def test_weight_scale_is_a_power_of_two_in_range():
    E = _emu()
    for amax in (1e-3, 0.024, 0.1, 0.7, 3.0):
        s = E.pow2_scale(amax, 2.0 ** 14)
        assert 2.0 ** 13 <= amax * s < 2.0 ** 14
        assert float(torch.tensor(s).log2()) == round(float(torch.tensor(s).log2()))


def test_nap_factor_forms_give_the_same_mahalanobis_score():
    """engine.nap_fit_from_stats on the CPU (torch.linalg.eigh / qr work on CPU tensors): the eigenvector form, the triangular
    whitening factor and the hybrid of both (triangular above the rounding-noise floor, eigenvector rows below) are the same
    quadratic form  sum_j ((d - mu) . v_j)^2 / var_j  -- utils/metric.py:220-222 -- and the hybrid's triangular block is upper
    triangular (what mmad_nap_set_structure promises the kernels)."""
    import torch
    from icra2021_multimodal_ad_b200.engine import nap_fit_from_stats
    g = torch.Generator().manual_seed(3)
    n, d, r = 400, 48, 40                                  # 8 directions at the noise floor (rank-deficient like SURVEY F5)
    basis = torch.randn(d, r, generator=g, dtype=torch.float64)
    spectrum = torch.logspace(0, -3, r, dtype=torch.float64)
    x = (torch.randn(n, r, generator=g, dtype=torch.float64) * spectrum) @ basis.t() + 1e-9 * torch.randn(n, d, generator=g, dtype=torch.float64)
    mu = x.mean(0)
    xc = x - mu
    gram = xc.t() @ xc
    test = (torch.randn(50, r, generator=g, dtype=torch.float64) * spectrum) @ basis.t() + 1e-9 * torch.randn(50, d, generator=g, dtype=torch.float64)
    scores = {}
    for factor in ("eigen", "triangular", "hybrid"):
        fit = nap_fit_from_stats(mu.float(), gram, n, factor=factor, tau=1e-5)
        vt, var = fit["vt"].double(), fit["var"].double()
        rot = (test - mu) @ vt.t()
        scores[factor] = ((rot * rot) / var).mean(1)
        t = fit["tri_rows"]
        assert (factor == "eigen") == (t == 0) and (factor != "triangular" or t == min(n, d))
        if t:
            assert torch.equal(torch.tril(fit["vt"][:t], diagonal=-1), torch.zeros_like(fit["vt"][:t]))
            assert float(fit["vt"][:t].abs().amax(1).min()) > 0.99          # rows normalised to unit max
    assert 0 < nap_fit_from_stats(mu.float(), gram, n, factor="hybrid", tau=1e-5)["tri_rows"] <= r
    # fp32 storage of the factor rows limits the agreement; the noise-floor directions dominate the absolute score
    assert float(((scores["hybrid"] - scores["eigen"]).abs() / scores["eigen"]).max()) < 1e-3
    strong = nap_fit_from_stats(mu.float(), gram[:0, :0] if False else gram, n, factor="triangular")
    assert strong["vt"].shape == (min(n, d), d)
