"""novelty_detection.py mirror (NoveltyDetecter.train / .test_arrays / .score_fast) end to end on synthetic data,
and the multi-GPU paths when the box has two devices (run with ``gpurun --gpus 2``; skipped otherwise)."""
import argparse
import os
import socket

import numpy as np
import pytest
import torch

from icra2021_multimodal_ad_b200.utils.synth import synth_windows

pytestmark = pytest.mark.gpu


def _cfg(D, **kw):
    return argparse.Namespace(input_size=D, btl_size=20, n_layers=3, gpu_id=0, n_epochs=3, batch_size=256, verbose=0,
                              start_layer_index=0, end_layer_index=-1, unimodal_normal=False, target_class=1,
                              train_diffs=None, **kw)


def test_train_then_test_like_the_reference_driver():
    from icra2021_multimodal_ad_b200.model_builder import get_model
    from icra2021_multimodal_ad_b200.novelty_detection import NoveltyDetecter
    D = 128
    cfg = _cfg(D, precision="f16x3")
    model = get_model(cfg)
    xtr, _ = synth_windows(2048, D, 1, anomaly_rate=0.0)
    xva, _ = synth_windows(512, D, 2, anomaly_rate=0.0)
    xte, yte = synth_windows(600, D, 3, anomaly_rate=0.2)
    loader = lambda x: [(x[i:i + 256], None) for i in range(0, len(x), 256)]  # noqa: E731
    det = NoveltyDetecter(cfg)
    tr_hist, va_hist, _, model = det.train(model, loader(xtr), loader(xva))
    assert len(tr_hist) == 3 and len(va_hist) == 3 and tr_hist[-1] < tr_hist[0]
    base, sap, nap, df = det.test_arrays(model, xtr, xva, xte, yte.numpy().astype(int))
    for auroc, aupr in (base, sap, nap):
        assert 0.0 <= float(auroc) <= 1.0 and 0.0 <= float(aupr) <= 1.0
    assert float(nap[0]) > 0.55                     # NAP separates the synthetic anomalies (SAP does not: they shrink a slice)
    assert list(df.columns)[:3] == ["base_auroc", "sap_auroc", "nap_auroc"]
    fast = det.score_fast(model, xtr, xva, xte, yte.numpy().astype(int))
    assert abs(float(fast["sap"]["auroc"]) - float(sap[0])) < 1e-6      # same scores -> same curve
    assert abs(float(fast["base"]["auroc"]) - float(base[0])) < 1e-6
    # NAP-fit checkpoint: the compact (mu, factor, var, mu2, N) artefact replaces the raw train diffs --
    # a second detector scores from it without the train set and gets the same NAP scores
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        cfg.nap_fit = os.path.join(tmp, "nap_fit.pt")
        a = det.score_fast(model, xtr, xva, xte, yte.numpy().astype(int))
        raw_bytes = len(xtr) * model.engine().concat_width(0, cfg.n_layers + 1) * 4
        assert os.path.getsize(cfg.nap_fit) < raw_bytes          # smaller than torch.save(train_diffs) (utils/metric.py:205)
        model.engine().load_state_dict(model.state_dict())        # drops the installed fit
        assert model.engine().nap_range is None
        b = NoveltyDetecter(cfg).score_fast(model, None, xva, xte, yte.numpy().astype(int))
        assert torch.allclose(a["nap"]["score"], b["nap"]["score"], rtol=1e-5, atol=0)
        assert abs(a["nap"]["auroc"] - b["nap"]["auroc"]) < 1e-6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from icra2021_multimodal_ad_b200 import parallel as P
        from icra2021_multimodal_ad_b200.model_builder import get_model
        from icra2021_multimodal_ad_b200.optim import Adam
        from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict
        D, btl, nl = 128, 100, 5
        cfg = argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=rank, precision="fp32")
        sd = synth_state_dict(D, btl, nl, 5)
        x, _ = synth_windows(96, D, 9, anomaly_rate=0.0)
        model = get_model(cfg)
        model.load_state_dict(sd)
        opt = Adam(model.parameters(), lr=1e-3)
        lo, hi = P.shard_range(len(x), rank, world)
        losses = [P.data_parallel_step(model, opt, x[lo:hi]) for _ in range(3)]
        trained = {k: v.cpu() for k, v in model.state_dict().items()}
        # the NVLink peer-memory exchange behind the BatchNorm statistics (csrc/peer.cu): against NCCL, every size class,
        # many back-to-back exchanges (both buffer sets, sequence numbers), identical bits on both ranks
        from icra2021_multimodal_ad_b200 import train as T
        from icra2021_multimodal_ad_b200._lib import check, lib
        st = T.train_state(model)
        peer_ok = bool(getattr(st, "peer", False))
        peer_err = 0.0
        if peer_ok:
            h = model.handle_engine()._h
            g = torch.Generator(device="cuda").manual_seed(100 + rank)
            for it in range(60):
                n = (1, 7, 100, 2816, 4096)[it % 5]
                v = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * 10 ** (it % 7 - 3)
                want = v.clone()
                dist.all_reduce(want)
                check(lib().mmad_peer_allreduce_f64(h, v.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
                peer_err = max(peer_err, float(((v - want).abs() / want.abs().clamp_min(1e-300)).max()))
                other = v.clone()
                dist.broadcast(other, src=0)
                assert torch.equal(other, v)            # bit-identical on every rank
        # sharded scoring + NAP fit over both ranks (fresh identical weights, well-conditioned selection [0:1])
        model.load_state_dict(sd)
        eng = model.eval().engine()
        xt, _ = synth_windows(700, D, 10, anomaly_rate=0.0)
        a, b = P.shard_range(len(xt), rank, world)
        eng.nap_fit(xt[a:b].cuda(), 0, 1)
        xs, _ = synth_windows(101, D, 11)
        sc = P.score_sharded(model, xs, 0, 1, nap=True)
        if rank == 0:
            q.put(dict(losses=losses, sd=trained, peer_ok=peer_ok, peer_err=peer_err,
                       nap=sc["nap"].cpu(), sap=sc["sap"].cpu()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_data_parallel_equals_one_gpu_on_the_concatenated_batch():
    import torch.multiprocessing as mp
    from icra2021_multimodal_ad_b200.model_builder import get_model
    from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
    from icra2021_multimodal_ad_b200.optim import Adam
    from icra2021_multimodal_ad_b200.reconstruction_aggregation import get_scores
    from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict
    import types
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got["peer_ok"], "the NVLink peer-memory exchange did not come up on this box"
    assert got["peer_err"] < 1e-14
    D, btl, nl = 128, 100, 5
    cfg = argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision="fp32")
    x, _ = synth_windows(96, D, 9, anomaly_rate=0.0)
    model = get_model(cfg)
    model.load_state_dict(synth_state_dict(D, btl, nl, 5))
    eng = types.SimpleNamespace(model=model, optimizer=Adam(model.parameters(), lr=1e-3), config=cfg)
    losses = [AutoEncoder.step(eng, (x, None))[0] for _ in range(3)]
    np.testing.assert_allclose(got["losses"], losses, rtol=2e-5)
    # Adam moves an element whose true gradient is zero (bias of an always-on LeakyReLU column feeding
    # BatchNorm) by +-lr per step on rounding noise -- in the reference too (tests/test_oracle_golden.py) -- so:
    # most elements (>70% even of the worst bias vector) agree to 5e-5 and none is further apart than the 3 steps allow
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            diff = (v.cpu() - got["sd"][k]).abs()
            assert (diff < 5e-5).float().mean().item() > 0.7, k
            assert diff.max().item() < 3 * 2.1e-3, k
    model.load_state_dict(synth_state_dict(D, btl, nl, 5))
    e1 = model.eval().engine()
    xt, _ = synth_windows(700, D, 10, anomaly_rate=0.0)
    e1.nap_fit(xt.cuda(), 0, 1, distributed=False)
    xs, _ = synth_windows(101, D, 11)
    sc = get_scores(xs, model, 0, 1, nap=True)
    np.testing.assert_allclose(got["sap"].numpy(), sc["sap"].cpu().numpy(), rtol=1e-6)     # same kernels, same rows
    np.testing.assert_allclose(got["nap"].numpy(), sc["nap"].cpu().numpy(), rtol=1e-3)     # fit from 2 shards vs 1
