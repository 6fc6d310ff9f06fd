"""GPU parity of the fused scoring path against the oracle and the golden fixtures.
Tolerances: activations/diffs 1e-5 relative to the matrix max, base/SAP
scores 1e-4 relative per sample, NAP 1e-4 on well-conditioned (single-layer) selections."""
import argparse

import numpy as np
import pytest
import torch

from conftest import load_golden
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

pytestmark = pytest.mark.gpu

PRECISIONS = ["fp32", "f16x3"]


def _rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _model(D, btl, nl, seed, precision="fp32"):
    from icra2021_multimodal_ad_b200.model_builder import get_model
    m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=precision))
    m.load_state_dict(synth_state_dict(D, btl, nl, seed))
    return m.eval()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", ["score_D64.pt", "score_D128_l3.pt", "score_D1728.pt"])
def test_scores_match_reference_golden(name, precision):
    from icra2021_multimodal_ad_b200.reconstruction_aggregation import get_diffs, get_scores
    g = load_golden(name)
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    m = _model(D, btl, nl, seed, precision)
    xte, _ = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
    r = g["xhat"].shape[0]
    with torch.no_grad():
        xhat = m(xte.cuda()).cpu()
        diffs = get_diffs(xte.numpy(), m, batch_size=g["bs"])
        loss = float(m.get_loss_value(xte.cuda(), xte.cuda()))
    assert _rel_max(xhat[:r], g["xhat"]) < 1e-5
    assert len(diffs) == nl + 1
    for d, dg in zip(diffs, g["diffs_te"]):
        assert d.dtype == np.float32 and d.shape[0] == g["n_te"]
        assert _rel_max(d[:r], dg) < 2e-5
    assert abs(loss - g["loss_sum_te"]) / g["loss_sum_te"] < 1e-5
    for sel, ent in g["sap"].items():
        lo, hi = sel.split(":")
        lo, hi = int(lo), (None if hi == "None" else int(hi))
        sc = get_scores(xte, m, lo, hi)
        np.testing.assert_allclose(sc["sap"].cpu().numpy(), ent["score"].numpy(), rtol=1e-4)
        np.testing.assert_allclose(sc["base"].cpu().numpy(), g["base"]["score"].numpy(), rtol=1e-4)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_against_oracle_ragged_and_edges(precision):
    from oracle import rapp_oracle as RO
    D, btl, nl, seed = 128, 100, 5, 77
    sd = synth_state_dict(D, btl, nl, seed)
    m = _model(D, btl, nl, seed, precision)
    eng = m.engine()
    for n in (1, 2, 127, 129, 300):
        x, _ = synth_windows(n, D, 500 + n)
        ref = RO.get_diffs(x, sd, batch_size=64)
        out = eng.score(x.cuda(), 0, nl + 1, diffs=True)
        cat = np.concatenate(ref, axis=1)
        assert _rel_max(out["diffs"].cpu().numpy(), cat) < 2e-5
        np.testing.assert_allclose(out["sap"].cpu().numpy(), RO.sap_score(ref), rtol=1e-4)
        np.testing.assert_allclose(out["base"].cpu().numpy(), RO.recon_score(ref[0]), rtol=1e-4)
        for lo, hi in ((1, 3), (2, 6), (5, 6), (0, 1)):
            o2 = eng.score(x.cuda(), lo, hi, base=False)
            np.testing.assert_allclose(o2["sap"].cpu().numpy(), RO.sap_score(ref, lo, hi), rtol=1e-4)
    # empty input
    out = eng.score(torch.empty(0, D, device="cuda"), 0, nl + 1)
    assert out["sap"].numel() == 0
    # strided rows (a column slice of a wider matrix)
    wide = torch.rand(100, D + 7, device="cuda")      # > 64 rows: both calls take the same kernel family
    o1 = eng.score(wide[:, :D], 0, nl + 1)["sap"]
    o2 = eng.score(wide[:, :D].contiguous(), 0, nl + 1)["sap"]
    if precision in ("fp32", "f16x3"):      # aligned rows of a model this narrow take the fused whole-chain kernels, the strided view the per-layer ones
        rel = float(((o1 - o2).abs() / o2.abs()).max())
        assert rel < (5e-5 if precision == "f16x3" else 1e-5), rel
    else:
        assert torch.equal(o1, o2)
    # determinism
    x, _ = synth_windows(300, D, 9)
    a = eng.score(x.cuda())["sap"]
    b = eng.score(x.cuda())["sap"]
    assert torch.equal(a, b)


def test_layer_by_layer_api_matches_fused_chain():
    """reconstruction_aggregation.py:25-27 iterates model.encoder.layer_list; the stand-alone
    FCLayer op must agree with the fused chain."""
    from oracle import rapp_oracle as RO
    D, btl, nl, seed = 64, 100, 5, 3
    m = _model(D, btl, nl, seed)
    sd = synth_state_dict(D, btl, nl, seed)
    x, _ = synth_windows(70, D, 1)
    h = x.cuda()
    ref = x
    for layer, L in zip(m.encoder.layer_list, RO.module_layers(sd, "encoder")):
        h = layer(h)
        ref = RO.fc_layer_eval(ref, L)
        assert _rel_max(h.cpu(), ref) < 1e-5
    z = m.encode(x.cuda())
    assert _rel_max(z.cpu(), ref) < 1e-5
    xh = m.decode(z)
    assert _rel_max(xh.cpu(), RO.ae_forward_eval(x, sd)) < 1e-5


@pytest.mark.parametrize("factor", ["eigen", "triangular"])
@pytest.mark.parametrize("precision", PRECISIONS)
def test_nap_single_layers_match_reference(precision, factor):
    g = load_golden("score_D64.pt")
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    m = _model(D, btl, nl, seed, precision)
    eng = m.engine()
    xtr, _ = synth_windows(g["n_tr"], D, seed + 1, anomaly_rate=0.0)
    xte, _ = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
    for sel in ("0:1", "1:2"):
        lo, hi = map(int, sel.split(":"))
        eng.nap_fit(xtr.cuda(), lo, hi, distributed=False, factor=factor)
        s = eng.score(xte.cuda(), lo, hi, base=False, sap=False, nap=True)["nap"].cpu().numpy()
        np.testing.assert_allclose(s, g["nap"][sel]["score"].numpy(), rtol=1e-3)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_nap_triangular_equals_eigen_at_headline_width(precision):
    """D = 1728, layer selection [0:1] (well conditioned, D' = 1728 > 1 tile row): the triangular whitening
    factor (half the products) and the eigenvector rotation give the same scores to 1e-4."""
    D, btl, nl, seed = 1728, 100, 5, 31
    m = _model(D, btl, nl, seed, precision)
    eng = m.engine()
    xtr, _ = synth_windows(4096, D, seed + 1, anomaly_rate=0.0)
    xte, _ = synth_windows(300, D, seed + 3, anomaly_rate=0.15)
    out = {}
    for factor in ("eigen", "triangular"):
        eng.nap_fit(xtr.cuda(), 0, 1, distributed=False, factor=factor)
        out[factor] = eng.score(xte.cuda(), 0, 1, base=False, sap=False, nap=True)["nap"].cpu().numpy()
    np.testing.assert_allclose(out["triangular"], out["eigen"], rtol=1e-4)


def test_f16_single_pass_stated_tolerance():
    """MMAD_PREC_F16 (one tensor-core pass on the fp16 hi parts): separately stated tolerance --
    per-sample SAP/base within 2e-2 relative, median within 3e-3 (precision modes)."""
    from icra2021_multimodal_ad_b200.reconstruction_aggregation import get_scores
    g = load_golden("score_D1728.pt")
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    m = _model(D, btl, nl, seed, "f16")
    xte, _ = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
    sc = get_scores(xte, m, 0, 7)
    ref = g["sap"]["0:7"]["score"].numpy()
    rel = np.abs(sc["sap"].cpu().numpy() - ref) / ref
    assert rel.max() < 2e-2 and np.median(rel) < 3e-3, (rel.max(), np.median(rel))


def test_tensor_core_gemm_tile_edges():
    """Shapes that exercise partial M/N/K tiles of the 128x256x64 tcgen05 kernel against the fp32 kernel."""
    for D, btl, nl in ((1728, 100, 5), (300, 17, 2), (2048, 100, 5), (93, 10, 3)):
        sd = synth_state_dict(D, btl, nl, 123)
        from icra2021_multimodal_ad_b200.model_builder import get_model
        outs = {}
        for prec in ("fp32", "f16x3"):
            m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=prec)).eval()
            m.load_state_dict(sd)
            for n in (1, 130, 700):
                x, _ = synth_windows(n, D, 40 + n)
                o = m.engine().score(x.cuda(), 0, nl + 1, diffs=True)
                outs[(prec, n)] = {k: v.cpu().numpy() for k, v in o.items()}
        for n in (1, 130, 700):
            a, b = outs[("fp32", n)], outs[("f16x3", n)]
            assert _rel_max(b["diffs"], a["diffs"]) < 2e-5, (D, n)
            np.testing.assert_allclose(b["sap"], a["sap"], rtol=5e-5)
            np.testing.assert_allclose(b["base"], a["base"], rtol=5e-5)


@pytest.mark.parametrize("D,btl,nl,n", [(1728, 100, 5, 2048 + 300), (300, 17, 2, 4096 + 129), (93, 10, 3, 2500)])
def test_cta_pair_kernel_matches_fp32_on_tall_chunks(D, btl, nl, n):
    """Chunks of >= 2048 rows run on the cta_group::2 kernel (gemm_tc2.cu): 256-row pair tiles with a ragged last
    tile (n not a multiple of 256 nor 128), partial N tiles, and the triangular NAP factor, against the fp32
    CUDA-core kernel on the same rows."""
    from icra2021_multimodal_ad_b200.model_builder import get_model
    sd = synth_state_dict(D, btl, nl, 77)
    x, _ = synth_windows(n, D, 5)
    xtr, _ = synth_windows(max(2 * D, 700), D, 6, anomaly_rate=0.0)
    outs = {}
    for prec in ("fp32", "f16x3"):
        m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=prec)).eval()
        m.load_state_dict(sd)
        eng = m.engine()
        eng.nap_fit(xtr.cuda(), 0, 1, distributed=False)
        o = eng.score(x.cuda(), 0, 1, nap=True, diffs=True)
        o2 = eng.score(x.cuda(), 0, nl + 1)
        outs[prec] = {k: v.cpu().numpy() for k, v in o.items()}
        outs[prec]["sap_all"] = o2["sap"].cpu().numpy()
    a, b = outs["fp32"], outs["f16x3"]
    assert _rel_max(b["diffs"], a["diffs"]) < 2e-5
    np.testing.assert_allclose(b["sap_all"], a["sap_all"], rtol=5e-5)
    np.testing.assert_allclose(b["base"], a["base"], rtol=5e-5)
    np.testing.assert_allclose(b["nap"], a["nap"], rtol=1e-3)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("n", [1, 8, 10, 33, 64])
def test_streaming_batches_take_the_fp32_weight_streaming_path(n, precision):
    """realtime_tester-style calls (test_file/realtime_tester.py:291-309): up to 16 windows run on exact-fp32 kernels
    whatever the handle's precision (f16x3, 17+ rows: split-K tensor-core layers + stand-alone epilogue, deterministic);
    device and host entry points, base/SAP/NAP and diffs against the oracle."""
    from oracle import rapp_oracle as RO
    D, btl, nl, seed = 1728, 100, 5, 31
    sd = synth_state_dict(D, btl, nl, seed)
    m = _model(D, btl, nl, seed, precision)
    eng = m.engine()
    xtr, _ = synth_windows(2000, D, seed + 1, anomaly_rate=0.0)
    eng.nap_fit(xtr.cuda(), 0, 1, distributed=False)
    x, _ = synth_windows(n, D, 100 + n)
    ref = RO.get_diffs(x, sd)
    o = eng.score(x.cuda(), 0, nl + 1, diffs=True)
    got = o["diffs"].cpu().numpy()
    off = 0
    tol = 1e-5 if n <= 32 else 2e-5
    for d in ref:
        assert _rel_max(got[:, off:off + d.shape[1]], d) < tol
        off += d.shape[1]
    np.testing.assert_allclose(o["sap"].cpu().numpy(), RO.sap_score(ref), rtol=5 * tol)
    np.testing.assert_allclose(o["base"].cpu().numpy(), RO.recon_score(ref[0]), rtol=5 * tol)
    big, _ = synth_windows(300, D, 100 + n)          # same first rows through the tensor-core path
    big[:n] = x
    nap_small = eng.score(x.cuda(), 0, 1, base=False, sap=False, nap=True)["nap"].cpu().numpy()
    nap_big = eng.score(big.cuda(), 0, 1, base=False, sap=False, nap=True)["nap"].cpu().numpy()[:n]
    np.testing.assert_allclose(nap_small, nap_big, rtol=1e-3)
    # host entry point: <= 64 windows without NAP run in ONE cooperative launch of exact-fp32 kernels (stream.cu), another
    # summation order than the per-layer kernels: compared with the oracle at the fp32 bar, deterministic call to call
    h = eng.score_host(x.numpy(), 0, nl + 1, base=True, sap=True, nap=False)
    np.testing.assert_allclose(h["sap"], RO.sap_score(ref), rtol=5e-5)
    np.testing.assert_allclose(h["base"], RO.recon_score(ref[0]), rtol=5e-5)
    h2 = eng.score_host(x.numpy(), 0, nl + 1, base=True, sap=True, nap=False)
    np.testing.assert_array_equal(h2["sap"], h["sap"])
    buf = eng.stream_input()                       # zero-copy variant: windows written into the pinned input buffer
    buf[:n] = x.numpy()
    h3 = eng.score_host(buf[:n], 0, nl + 1, base=True, sap=True, nap=False)
    np.testing.assert_array_equal(h3["sap"], h["sap"])
    hn = eng.score_host(x.numpy(), 0, 1, base=True, sap=True, nap=True)            # with NAP: graph-replay path
    np.testing.assert_allclose(hn["nap"], nap_small, rtol=1e-5)
    np.testing.assert_allclose(hn["sap"], RO.sap_score(ref, 0, 1), rtol=5e-5)


def test_bulk_size_properties():
    """Bulk size (300 000 windows of the headline model: several device chunks, partial last chunk): properties that
    do not need an oracle at that size -- rows are scored independently of their neighbours and of the chunking
    (bit-identical under permutation and under splitting the call) -- plus the oracle on a random subset."""
    from oracle import rapp_oracle as RO
    D, btl, nl, seed = 1728, 100, 5, 31
    sd = synth_state_dict(D, btl, nl, seed)
    m = _model(D, btl, nl, seed, "f16x3")
    eng = m.engine()
    xtr, _ = synth_windows(2000, D, seed + 1, anomaly_rate=0.0)
    eng.nap_fit(xtr.cuda(), 0, 1, distributed=False)
    n = 300_000
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(n, D, device="cuda", generator=g)
    full = eng.score(x, 0, 1, nap=True)
    perm = torch.randperm(n, device="cuda", generator=g)
    shuffled = eng.score(x[perm], 0, 1, nap=True)
    for k in ("base", "sap", "nap"):
        assert torch.equal(shuffled[k], full[k][perm]), k
    cut = 123_457
    a, b = eng.score(x[:cut], 0, 1, nap=True), eng.score(x[cut:], 0, 1, nap=True)
    for k in ("base", "sap", "nap"):
        assert torch.equal(torch.cat([a[k], b[k]]), full[k]), k
    idx = torch.randint(0, n, (48,), device="cuda", generator=g)
    ref = RO.get_diffs(x[idx].cpu(), sd)
    np.testing.assert_allclose(full["base"][idx].cpu().numpy(), RO.recon_score(ref[0]), rtol=1e-4)
    sap_all = eng.score(x[idx.sort().values], 0, nl + 1)["sap"]           # 48 rows: weight-streaming path
    big = eng.score(x, 0, nl + 1)["sap"][idx.sort().values]
    np.testing.assert_allclose(big.cpu().numpy(), sap_all.cpu().numpy(), rtol=1e-4)
    assert torch.isfinite(full["nap"]).all()
