"""GPU parity of the fused multimodal feature extractor (utils/data_loaders.py:152-229, 601-674, 703-731) against the
golden outputs of the unmodified reference classes and against the oracle; and the extractor feeding the scorer."""
import argparse

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_fused_extractor_matches_reference_golden():
    from icra2021_multimodal_ad_b200.utils.data_loaders import HSR_Net, Multisensory_module, norm_vec
    g = load_golden("features.pt")
    B = g["r"].shape[0]
    cfg = argparse.Namespace(batch_size=B, slicing_size=B, gpu_id=0)
    net = Multisensory_module(cfg).cuda()
    assert list(net.state_dict().keys()) == list(g["sd"].keys())
    net.load_state_dict(g["sd"])
    out = net(g["r"].cuda(), g["d"].cuda(), g["t"].cuda(), g["m"].cuda())
    assert tuple(out.shape) == (B, 27, 8, 8) and out.view(B, -1).shape[1] == 1728
    assert _rel(out, g["fused"]) < 1e-5
    hsr = HSR_Net(True, cfg).cuda()
    hsr.load_state_dict(g["sd"])
    assert _rel(hsr(g["r"].cuda(), None, None, None, None), g["rgb_only"]) < 1e-5
    assert _rel(hsr(None, g["d"].cuda(), None, None, None), g["depth_only"]) < 1e-5
    with pytest.raises(NotImplementedError):
        hsr(None, None, g["r"].cuda(), None, None)
    assert _rel(norm_vec(g["raw_r"], [0, 255]), g["normed"]["r"]) < 1e-6
    assert _rel(norm_vec(g["raw_m"]), g["normed"]["m"]) < 1e-6


@pytest.mark.parametrize("B", [1, 10, 257])
def test_extractor_against_oracle_and_into_the_scorer(B):
    """Raw sensor ranges through HsrDataset (normalisation folded into the kernel) == oracle; the 1728-d vectors go
    straight into the fused scorer like test_file/realtime_tester.py:291-304."""
    from icra2021_multimodal_ad_b200.model_builder import get_model
    from icra2021_multimodal_ad_b200.reconstruction_aggregation import get_scores
    from icra2021_multimodal_ad_b200.utils import data_loaders as DL
    from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict
    from oracle import feature_oracle as FO
    from oracle import rapp_oracle as RO
    g = torch.Generator().manual_seed(B)
    hand = torch.rand(B, 3 * 32 * 32, generator=g) * 255
    depth = torch.rand(B, 32 * 32, generator=g) * 255
    force = torch.rand(B, generator=g) * 400
    mic = torch.randn(B, 13, generator=g) * 20
    cfg = argparse.Namespace(batch_size=B, gpu_id=0)
    torch.manual_seed(5)
    fusion = DL.HsrDataset(cfg, force.numpy(), hand.numpy(), depth.numpy(), mic.numpy())
    assert tuple(fusion.shape) == (B, 1728)
    torch.manual_seed(5)
    ref_net = DL.Multisensory_module(cfg)            # same initialisation as inside HsrDataset
    sd = {k: v.detach().cpu() for k, v in ref_net.state_dict().items()}
    ref = FO.multisensory_forward(sd, FO.norm_vec(hand.view(B, 1, 3, 32, 32), [0, 255]),
                                  FO.norm_vec(depth.view(B, 1, 1, 32, 32), [0, 255]), FO.norm_vec(force, [0, 400]),
                                  FO.norm_vec(mic.view(B, 1, 1, 13))).reshape(B, -1)
    assert _rel(fusion, ref) < 2e-5
    m = get_model(argparse.Namespace(input_size=1728, btl_size=100, n_layers=5, gpu_id=0, precision="f16x3")).eval()
    ae_sd = synth_state_dict(1728, 100, 5, 2)
    m.load_state_dict(ae_sd)
    sc = get_scores(fusion, m)
    want = RO.sap_score(RO.get_diffs(ref, ae_sd))
    np.testing.assert_allclose(sc["sap"].cpu().numpy(), want, rtol=2e-4)
