#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no golden vectors; these outputs of its own
functions are what pins ``oracle/`` and, through it, the CUDA path.  Two import shims
are needed and no reference file is modified: ``collections.Iterable``
(models/abstract_model.py:25) and a stub ``matplotlib.pyplot`` (utils/metric.py:21).
"""
import argparse
import collections
import collections.abc
import contextlib
import io
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows  # noqa: E402

collections.Iterable = collections.abc.Iterable
_m, _p = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
_m.pyplot = _p
sys.modules["matplotlib"] = _m
sys.modules["matplotlib.pyplot"] = _p
sys.path.insert(0, "/root/reference")
from model_builder import ae_wrapper  # noqa: E402
from models.auto_encoder import AutoEncoder  # noqa: E402
from modules import FCModule  # noqa: E402
from reconstruction_aggregation import get_diffs  # noqa: E402
import utils.metric as M  # noqa: E402

warnings.filterwarnings("ignore")
torch.set_num_threads(8)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ref_model(D, btl, nl, seed):
    cfg = argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=-1, train_diffs="/tmp/_td.pt")
    model = ae_wrapper(cfg)
    model.load_state_dict(synth_state_dict(D, btl, nl, seed))
    model.eval()
    return model, cfg


def scoring_case(name, D, btl, nl, seed, n_tr, n_va, n_te, bs, selections, keep_rows=None):
    model, cfg = ref_model(D, btl, nl, seed)
    xtr, _ = synth_windows(n_tr, D, seed + 1, anomaly_rate=0.0)
    xva, _ = synth_windows(n_va, D, seed + 2, anomaly_rate=0.0)
    xte, yte = synth_windows(n_te, D, seed + 3, anomaly_rate=0.15)
    y = yte.numpy().astype(bool)
    with torch.no_grad():
        dtr = get_diffs(xtr, model, batch_size=bs)
        dva = get_diffs(xva, model)
        dte = get_diffs(xte.numpy(), model, batch_size=bs)   # ndarray input path (line 8-9)
        xhat = model(xte).numpy()
    out = dict(D=D, btl=btl, n_layers=nl, seed=seed, n_tr=n_tr, n_va=n_va, n_te=n_te, bs=bs)
    r = keep_rows or n_te
    out["xhat"] = torch.from_numpy(xhat[:r].copy())
    out["diffs_te"] = [torch.from_numpy(d[:r].copy()) for d in dte]
    out["loss_sum_te"] = float(model.get_loss_value(xte, xte))
    base = quiet(M.get_recon_loss, dva[0], dte[0], y, f1_quantiles=[.90])
    out["base"] = dict(score=torch.from_numpy(base[0]), metrics=[float(v) for v in base[1:]])
    out["sap"] = {}
    out["nap"] = {}
    for (lo, hi) in selections:
        sap = quiet(M.get_d_loss, dtr, dva, dte, y, start_layer_index=lo, end_layer_index=hi,
                    gpu_id=-1, norm_type=2, f1_quantiles=[.90])
        out["sap"][f"{lo}:{hi}"] = dict(score=torch.from_numpy(sap[0]), metrics=[float(v) for v in sap[1:]])
        nap = quiet(M.get_d_norm_loss, dtr, dva, dte, y, cfg, start_layer_index=lo, end_layer_index=hi,
                    gpu_id=-1, norm_type=2, f1_quantiles=[.90])
        out["nap"][f"{lo}:{hi}"] = dict(score=torch.from_numpy(np.asarray(nap[0])),
                                        metrics=[float(v) for v in nap[1:]])
    torch.save(out, os.path.join(HERE, name))
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def nap_full_rank_case(name, D, btl, nl, seed, n_tr, n_te):
    """All-layers NAP (novelty_detection.py:56-57,69-70) with N_tr >= D' so that K = D' like the benchmarked fit: only the
    per-window NAP scores and their metrics are kept."""
    model, cfg = ref_model(D, btl, nl, seed)
    xtr, _ = synth_windows(n_tr, D, seed + 1, anomaly_rate=0.0)
    xva, _ = synth_windows(256, D, seed + 2, anomaly_rate=0.0)
    xte, yte = synth_windows(n_te, D, seed + 3, anomaly_rate=0.15)
    y = yte.numpy().astype(bool)
    with torch.no_grad():
        dtr = get_diffs(xtr, model, batch_size=256)
        dva = get_diffs(xva, model)
        dte = get_diffs(xte, model, batch_size=256)
    nap = quiet(M.get_d_norm_loss, dtr, dva, dte, y, cfg, start_layer_index=0, end_layer_index=nl + 2,
                gpu_id=-1, norm_type=2, f1_quantiles=[.90])
    torch.save(dict(D=D, btl=btl, n_layers=nl, seed=seed, n_tr=n_tr, n_te=n_te,
                    nap=dict(score=torch.from_numpy(np.asarray(nap[0])), metrics=[float(v) for v in nap[1:]])),
               os.path.join(HERE, name))
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def train_case(name, D, btl, nl, seed, B, steps, full_state):
    model, cfg = ref_model(D, btl, nl, seed)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    eng = types.SimpleNamespace(model=model, optimizer=opt, config=cfg)
    losses = []
    states = []
    grads = []
    for s in range(steps):
        xb, _ = synth_windows(B, D, seed + 100 + s, anomaly_rate=0.0)
        loss, = AutoEncoder.step(eng, (xb, None))
        losses.append(loss)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        gr = {k: p.grad.clone() for k, p in model.named_parameters()}
        if not full_state:   # keep BN buffers, biases and a corner of every weight
            sd = {k: (v if v.dim() < 2 else v[:8, :8].clone()) for k, v in sd.items()}
            gr = {k: (v if v.dim() < 2 else v[:8, :8].clone()) for k, v in gr.items()}
        states.append(sd)
        grads.append(gr)
    xv, _ = synth_windows(B, D, seed + 999, anomaly_rate=0.0)
    vloss, = AutoEncoder.validate(eng, (xv, None))
    torch.save(dict(D=D, btl=btl, n_layers=nl, seed=seed, B=B, steps=steps, losses=losses, states=states, grads=grads,
                    full_state=full_state, valid_loss=vloss), os.path.join(HERE, name))
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def vib_case(name):
    D, out_sz, B, k = 64, 40, 12, 3
    torch.manual_seed(5)
    enc = FCModule(input_size=D, output_size=out_sz, hidden_sizes=[56, 48], use_batch_norm=True,
                   act="leakyrelu", last_act=None)
    enc.eval()
    x = torch.rand(B, D, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        torch.manual_seed(99)
        r = enc(x, distribution="normal", k=k)
        torch.manual_seed(99)
        eps = torch.randn(k, B, out_sz // 2)
        det = enc(x, distribution="normal", k=k, stochastic_inference=False)
        plain = enc(x)
    torch.save(dict(sd=enc.state_dict(), x=x, eps=eps, z=r["z"], mu=r["mu"], logvar=r["logvar"],
                    z_det=det["z"].clone(), plain=plain, k=k), os.path.join(HERE, name))
    print(name)


def feature_case(name):
    """utils/data_loaders.py: Multisensory_module / HSR_Net / norm_vec, run unmodified on CPU.  Extra shims for this
    module only: stub ``librosa`` (imported at the top of utils/data_loaders.py, used by the wav loader only) and a
    no-op ``.cuda()`` (the forward allocates its output with ``torch.Tensor().cuda(gpu_id)``)."""
    for mod in ("librosa", "librosa.display", "librosa.feature"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    tcuda, mcuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        import utils.data_loaders as DL
        B = 5
        g = torch.Generator().manual_seed(77)
        torch.manual_seed(3)
        cfg = argparse.Namespace(batch_size=B, slicing_size=B, gpu_id=0)
        net = DL.Multisensory_module(cfg)
        r = torch.rand(B, 1, 3, 32, 32, generator=g) * 2 - 1
        d = torch.rand(B, 1, 1, 32, 32, generator=g) * 2 - 1
        t = torch.rand(B, generator=g) * 2 - 1
        m = torch.rand(B, 1, 1, 13, generator=g) * 2 - 1
        with torch.no_grad():
            fused = quiet(net, r, d, t, m)
            hsr = DL.HSR_Net(True, cfg)
            hsr.load_state_dict(net.state_dict())
            rgb_only = quiet(hsr, r, None, None, None, None)
            depth_only = quiet(hsr, None, d, None, None, None)
            # HsrDataset normalisation (raw sensor ranges)
            raw_r = torch.rand(B, 3 * 32 * 32, generator=g) * 255
            raw_t = torch.rand(B, generator=g) * 400
            raw_m = torch.randn(B, 13, generator=g) * 30
            normed = dict(r=DL.norm_vec(raw_r, range_in=[0, 255]), t=DL.norm_vec(raw_t, range_in=[0, 400]), m=DL.norm_vec(raw_m))
        torch.save(dict(sd=net.state_dict(), r=r, d=d, t=t, m=m, fused=fused, rgb_only=rgb_only, depth_only=depth_only,
                        raw_r=raw_r, raw_t=raw_t, raw_m=raw_m, normed=normed), os.path.join(HERE, name))
        print(name, tuple(fused.shape))
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = tcuda, mcuda


def metric_cases(name):
    rng = np.random.default_rng(7)
    cases = []

    def add(tag, s, y, v=None):
        s = np.asarray(s, dtype=np.float32)
        y = np.asarray(y, dtype=bool)
        v = s if v is None else np.asarray(v, dtype=np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                roc = float(M.get_auc_roc(s, y))
                prc = float(M.get_auc_prc(s, y))
                f1, thr = M.get_f1_score(v, s, y)
                try:
                    p, r = quiet(M.get_confusion_matrix, s, y, thr)
                except Exception:
                    p, r = float("nan"), float("nan")
        cases.append(dict(tag=tag, score=[float(x) for x in s], label=[int(b) for b in y],
                          valid=[float(x) for x in v], auroc=roc, auprc=prc, f1=float(f1), thr=float(thr),
                          precision=float(p), recall=float(r)))

    add("doc", [.1, .4, .35, .8, .8, .2], [0, 0, 1, 1, 0, 1])
    add("all_neg", [.1, .4, .35, .8], [0, 0, 0, 0])
    add("all_pos", [.1, .4, .35, .8], [1, 1, 1, 1])
    add("all_tied", [.5] * 8, [0, 1, 0, 1, 1, 0, 0, 1])
    add("nan", [.1, float("nan"), .3, .4], [0, 1, 0, 1])
    add("inf", [.1, float("inf"), .3, .4], [0, 1, 0, 1])
    add("two", [.2, .7], [0, 1])
    for n in (7, 8, 9, 64, 127, 128, 129, 257, 1000, 4097):
        s = rng.random(n).astype(np.float32)
        y = rng.random(n) < 0.2
        y[0], y[1] = True, False
        add(f"rand{n}", s + 0.5 * y * rng.random(n).astype(np.float32), y, rng.random(max(n // 2, 3)))
    for n in (50, 600, 3000):   # heavy ties
        s = np.round(rng.random(n) * 10).astype(np.float32) / 10
        y = rng.random(n) < 0.3
        y[0], y[1] = True, False
        add(f"ties{n}", s, y, np.round(rng.random(n) * 10) / 10)
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(cases, f)
    print(name, len(cases), "cases")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nap_full":     # the slow case alone (SVD of 6144 x 5482 on the CPU)
        nap_full_rank_case("nap_D1728_full.pt", 1728, 100, 5, 31, n_tr=6144, n_te=384)
        sys.exit(0)
    # real per-modality dims (utils/data_loaders.py:16-29): force_torque 64, mic 128
    scoring_case("score_D64.pt", 64, 100, 5, 11, n_tr=700, n_va=90, n_te=130, bs=48,
                 selections=[(0, 7), (0, 1), (1, 2), (2, 5), (5, 6), (9, 3), (0, None)])
    scoring_case("score_D128_l3.pt", 128, 10, 3, 21, n_tr=600, n_va=64, n_te=100, bs=33,
                 selections=[(0, 5), (1, 3), (3, 4)])
    # headline dims; only the first rows of the big arrays are kept
    scoring_case("score_D1728.pt", 1728, 100, 5, 31, n_tr=2048, n_va=256, n_te=384, bs=256,
                 selections=[(0, 7), (0, 1), (1, 2), (5, 6)], keep_rows=6)
    train_case("train_D64.pt", 64, 100, 5, 41, B=32, steps=3, full_state=True)
    train_case("train_D1728.pt", 1728, 100, 5, 51, B=256, steps=2, full_state=False)
    vib_case("vib_D64.pt")
    feature_case("features.pt")
    metric_cases("metrics_golden.json")
    nap_full_rank_case("nap_D1728_full.pt", 1728, 100, 5, 31, n_tr=6144, n_te=384)
