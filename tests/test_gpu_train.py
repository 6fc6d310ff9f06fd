"""GPU parity of the fused training step (models/auto_encoder.py:57-77) against the golden fixtures
produced by the unmodified reference (AutoEncoder.step + torch.optim.Adam on CPU) and against the
oracle.  Tolerances follow tests/test_oracle_golden.py::test_train_step_matches_reference, which pins
the oracle on the same fixtures: loss 1e-5 relative at step 0, gradients 2e-4 of the tensor max,
post-Adam parameters 2e-5 absolute where the gradient is above rounding noise."""
import argparse
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

pytestmark = pytest.mark.gpu


def _grad_ok(mine, ref, full_scale, strict, key):
    """strict: every element within 2e-4 of the tensor max (the oracle's own bar against the golden).
    robust: used at the headline dims, where ~1e6 pre-activations per step make a LeakyReLU branch flip
    (|pre| ~ 1e-7) between two fp32 implementations likely; ONE flip moves every upstream gradient by
    ~2e-3 of its max and the hit row by ~2e-2 (measured on the reference algorithm itself by perturbing
    the weights by 6e-8: scripts/chaos_train_oracle.py; fp32 vs fp64: scripts/diag_train.py), so the bar is
    98% of the elements within 2e-2 of the max and a relative Frobenius error below 1e-1 (a few flips).
    The strict bar is enforced at the same widths with a small batch (no flip expected)."""
    err = (mine.double() - ref.double()).abs() / full_scale
    if strict:
        assert err.max().item() < 2e-4, key
    else:
        assert torch.quantile(err.flatten(), 0.98).item() < 2e-2, key
        assert (mine.double() - ref.double()).norm().item() <= 1e-1 * max(ref.double().norm().item(), 1e-30) + 1e-6 * full_scale, key


def _model(D, btl, nl, seed, **kw):
    from icra2021_multimodal_ad_b200.model_builder import get_model
    m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, **kw))
    return m


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("optimizer", ["torch", "mmad"])
@pytest.mark.parametrize("name", ["train_D64.pt", "train_D1728.pt"])
def test_step_matches_reference_golden(name, optimizer, precision):
    from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
    from icra2021_multimodal_ad_b200.optim import Adam
    g = load_golden(name)
    D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
    m = _model(D, btl, nl, seed, precision=precision)
    m.load_state_dict(synth_state_dict(D, btl, nl, seed))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3) if optimizer == "torch" else Adam(m.parameters(), lr=1e-3)
    eng = types.SimpleNamespace(model=m, optimizer=opt, config=argparse.Namespace(gpu_id=0))
    noisy = {}
    for s in range(g["steps"]):
        xb, _ = synth_windows(g["B"], D, seed + 100 + s, anomaly_rate=0.0)
        loss, = AutoEncoder.step(eng, (xb, None))            # the reference's step body, verbatim
        assert isinstance(loss, float)
        assert abs(loss - g["losses"][s]) / g["losses"][s] < (1e-5 if s == 0 else 2e-3)
        grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
        for k, gr in g["grads"][s].items():
            mine = grads[k]
            if not g["full_state"] and mine.dim() == 2:
                mine = mine[:8, :8]
            scale = max(grads[k].abs().max().item(), 1e-12)
            if s == 0:
                _grad_ok(mine, gr, scale, strict=D < 1000, key=k)
            noisy[k] = noisy.get(k, torch.zeros_like(gr, dtype=torch.bool)) | (gr.abs() < (1e-3 if D < 1000 else 3e-2) * scale)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        for k, v in g["states"][s].items():
            mine = sd[k]
            if not g["full_state"] and mine.dim() == 2:
                mine = mine[:8, :8]
            if k.endswith("num_batches_tracked"):
                assert int(mine) == int(v)
            elif k in noisy:
                ok = ~noisy[k]
                assert ((mine - v).abs() * ok).max().item() < (2e-5 if s == 0 else 2.5e-3), k
            else:
                assert (mine - v).abs().max().item() < (2e-6 if s == 0 else 1e-3), k
    xv, _ = synth_windows(g["B"], D, seed + 999, anomaly_rate=0.0)
    vloss, = AutoEncoder.validate(eng, (xv, None))            # eval engine re-packs the updated weights
    assert abs(vloss - g["valid_loss"]) / g["valid_loss"] < 5e-3


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("D,B", [(128, 7), (64, 1), (1728, 300), (1728, 24), (93, 33), (128, 2100)])
def test_gradients_match_oracle(D, B, precision):
    """Odd batch sizes / widths (ragged tiles; B = 2100 takes the CTA-pair forward kernel and split-K dW) against
    the oracle's manual backward.  Small cases are held
    to the strict bar on every draw; at D = 1728 a LeakyReLU branch flip in either implementation is a coin
    toss per draw (see _grad_ok), so three draws are taken: all must meet the robust bar and at least one
    (a flip-free one) the strict bar."""
    from oracle import rapp_oracle as RO
    btl, nl, seed = (100, 5, 3) if D != 93 else (10, 3, 4)
    sd = synth_state_dict(D, btl, nl, seed)
    m = _model(D, btl, nl, seed, precision=precision)
    small = B * D < 20000
    worst = []
    for xseed in ((77,) if small else (77, 78, 79)):
        x, _ = synth_windows(B, D, xseed, anomaly_rate=0.0)
        m.load_state_dict(sd)
        m.train()
        m.zero_grad()
        if B == 1:
            # torch raises for a single-row train-mode BatchNorm; ours is defined (var 0) -- just run it
            loss = m.get_loss_value(x.cuda(), x.cuda())
            loss.backward()
            assert np.isfinite(float(loss.detach()))
            return
        ref_loss, ref_grads, ref_bufs = RO.train_forward_backward(x, dict(sd))
        loss = m.get_loss_value(x.cuda(), x.cuda())
        loss.backward()
        assert abs(float(loss.detach()) - ref_loss) / ref_loss < 1e-5
        w = 0.0
        for k, p in m.named_parameters():
            gr = ref_grads[k]
            scale = max(gr.abs().max().item(), 1e-12)
            _grad_ok(p.grad.cpu(), gr, scale, strict=small, key=k)
            w = max(w, ((p.grad.cpu() - gr).abs().max() / scale).item())
        worst.append(w)
        sdm = m.state_dict()
        for k, v in ref_bufs.items():
            if k.endswith("num_batches_tracked"):
                assert int(sdm[k]) == int(v)
            else:
                assert (sdm[k].cpu() - v).abs().max().item() < 2e-6, k
    if B * D < 100000:
        assert min(worst) < 2e-4, worst


def test_zero_grad_set_to_none_false_does_not_double_count():
    D, btl, nl, seed = 64, 100, 5, 9
    x, _ = synth_windows(16, D, 5, anomaly_rate=0.0)
    m = _model(D, btl, nl, seed)
    m.train()
    m.get_loss_value(x.cuda(), x.cuda()).backward()
    g1 = [p.grad.clone() for p in m.parameters()]
    for p in m.parameters():
        p.grad.zero_()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    m.get_loss_value(x.cuda(), x.cuda()).backward()
    for a, p in zip(g1, m.parameters()):
        torch.testing.assert_close(p.grad, a, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
def test_vib_training_step_matches_autograd_restatement(precision):
    """VIB autoencoder (BASELINE configs[3]): reparameterisation pinned by the reference, KL
    parity-unpinned -- compared with the oracle's torch-autograd fp32 statement."""
    from oracle import rapp_oracle as RO
    D, btl, nl, seed, B, beta = 128, 20, 4, 13, 48, 0.5
    sd = synth_state_dict(D, btl, nl, seed, enc_out=2 * btl)
    x, _ = synth_windows(B, D, 21, anomaly_rate=0.0)
    eps = torch.randn(B, btl, generator=torch.Generator().manual_seed(99))
    ref_loss, ref_grads, ref_bufs = RO.vib_train_forward_backward(x, dict(sd), eps, beta)
    m = _model(D, btl, nl, seed, vib=True, beta_kl=beta, precision=precision)
    m.load_state_dict(sd)
    m.train()
    loss = m.get_loss_value(x.cuda(), x.cuda(), eps=eps.cuda())
    loss.backward()
    assert abs(float(loss) - ref_loss) / abs(ref_loss) < 1e-5
    for k, p in m.named_parameters():
        gr = ref_grads[k]
        scale = max(gr.abs().max().item(), 1e-12)
        assert (p.grad.cpu() - gr).abs().max().item() / scale < 2e-4, k
    sdm = m.state_dict()
    for k, v in ref_bufs.items():
        if not k.endswith("num_batches_tracked"):
            assert (sdm[k].cpu() - v).abs().max().item() < 2e-6, k


def test_training_reduces_loss_and_scoring_follows():
    """A few hundred fused steps: the loss falls, and the eval engine scores with the trained weights."""
    from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
    from icra2021_multimodal_ad_b200.optim import Adam
    from icra2021_multimodal_ad_b200.reconstruction_aggregation import get_scores
    from oracle import rapp_oracle as RO
    D, btl, nl = 128, 100, 5
    m = _model(D, btl, nl, 0)
    opt = Adam(m.parameters(), lr=1e-3)
    eng = types.SimpleNamespace(model=m, optimizer=opt, config=argparse.Namespace(gpu_id=0))
    xw, _ = synth_windows(2048, D, 7, anomaly_rate=0.0)
    xw = xw.cuda()
    losses = [AutoEncoder.step(eng, (xw[(i * 256) % 2048:(i * 256) % 2048 + 256], None))[0] for i in range(120)]
    assert losses[-1] < 0.5 * losses[0]
    xt, _ = synth_windows(300, D, 8)
    with torch.no_grad():
        sc = get_scores(xt, m)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = RO.get_diffs(xt, sd)
    np.testing.assert_allclose(sc["sap"].cpu().numpy(), RO.sap_score(ref), rtol=1e-4)
