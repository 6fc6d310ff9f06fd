"""Hybrid NAP factor: how the split point tau (triangular factor for singular values >= tau * sigma_1, eigenvector rows below)
trades numerical quality against tensor work.  (a) the reference golden at D = 1728, all layers, N_tr = 6144: error / rank
agreement against the fp64 value, AUROC; (b) bench.py's model (20 train steps, 65 536-row fit): triangular rows and ms/step.
python scripts/nap_tau_sweep.py"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from scipy.stats import spearmanr
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils import metric as M
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
from oracle import rapp_oracle as RO
import bench

g = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "nap_D1728_full.pt"), weights_only=False)
D, btl, nl, seed = g["D"], g["btl"], g["n_layers"], g["seed"]
sd = synth_state_dict(D, btl, nl, seed)
xtr, _ = synth_windows(g["n_tr"], D, seed + 1, anomaly_rate=0.0)
xte, yte = synth_windows(g["n_te"], D, seed + 3, anomaly_rate=0.15)
y = yte.numpy().astype(bool)
truth = RO.nap_score_fp64(RO.concat_diffs(RO.get_diffs(xtr, sd)), RO.concat_diffs(RO.get_diffs(xte, sd)))
ref = g["nap"]["score"].numpy().astype(np.float64)
print("reference: err %.3f rho %.4f auroc %.4f" % (np.median(np.abs(ref - truth) / truth), spearmanr(ref, truth).correlation, g["nap"]["metrics"][0]))
for prec in (() if os.environ.get("ONLY_C") else ("f16x3", "fp32")):
    m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=prec)).eval()
    m.load_state_dict(sd)
    eng = m.engine()
    for tau in (1e-2, 1e-3, 3e-4, 1e-4, 3e-5, 1e-5, 0.0):
        fit = eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False, factor="hybrid", tau=tau)
        new = eng.score(xte.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
        print("golden %s tau %.0e: triangular rows %d / %d  err %.3f rho %.4f auroc %.4f" % (
            prec, tau, fit["tri_rows"], fit["vt"].shape[0], np.median(np.abs(new - truth) / truth), spearmanr(new, truth).correlation,
            M.get_auc_roc(new.astype(np.float32), y)), flush=True)

# bench model
sdb = bench.trained_state_dict(0)
m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision="f16x3")).eval()
m.load_state_dict(sdb)
eng = m.engine()
xs, _ = synth_windows(8192, D, 1234, anomaly_rate=0.0)
xf = xs.repeat(8, 1)
xf = (xf + 1e-3 * torch.randn(xf.shape, generator=torch.Generator().manual_seed(0))).clamp_(0, 1).cuda()
xb, _ = synth_windows(8192, D, 1236)
xb = xb.repeat(10, 1)[:75776].contiguous().cuda()
for tau in (() if os.environ.get("ONLY_C") else (1e-2, 1e-3, 3e-4, 1e-4, 3e-5, 1e-5, 0.0)):
    fit = eng.nap_fit(xf, 0, nl + 1, distributed=False, factor="hybrid", tau=tau)
    for _ in range(3):
        eng.score(xb, 0, nl + 1, nap=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.score(xb, 0, nl + 1, nap=True)
    e1.record(); torch.cuda.synchronize()
    lam = None
    print("bench model tau %.0e: triangular rows %d / %d  %.2f ms/step" % (tau, fit["tri_rows"], fit["vt"].shape[0], e0.elapsed_time(e1) / 5), flush=True)

# (c) the protocol on bench.py's TRAINED model: oracle fp32 NAP (the reference's algorithm) and the fp64 value on the oracle's diffs
xtr2, _ = synth_windows(6144, D, 41, anomaly_rate=0.0)
xte2, yte2 = synth_windows(512, D, 43, anomaly_rate=0.15)
y2 = yte2.numpy().astype(bool)
dtr, dte = RO.get_diffs(xtr2, sdb), RO.get_diffs(xte2, sdb)
ctr, cte = RO.concat_diffs(dtr), RO.concat_diffs(dte)
truth2 = RO.nap_score_fp64(ctr, cte)
ref2 = RO.NapFit(ctr).score(cte).astype(np.float64)
print("trained model, oracle (reference algorithm, fp32): err %.3f rho %.4f auroc %.4f | fp64 auroc %.4f" % (
    np.median(np.abs(ref2 - truth2) / truth2), spearmanr(ref2, truth2).correlation, M.get_auc_roc(ref2.astype(np.float32), y2),
    M.get_auc_roc(truth2.astype(np.float32), y2)), flush=True)
for prec in ("f16x3", "fp32"):
    m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=prec)).eval()
    m.load_state_dict(sdb)
    eng = m.engine()
    for factor, tau in (("hybrid", 1e-3), ("hybrid", 1e-5), ("hybrid", 1e-6), ("triangular", 0.0), ("eigen", 0.0)):
        fit = eng.nap_fit(xtr2.cuda(), 0, nl + 1, distributed=False, factor=factor, tau=tau)
        new = eng.score(xte2.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
        print("trained %s %s tau %.0e: triangular rows %d / %d  err %.3f rho %.4f auroc %.4f" % (
            prec, factor, tau, fit["tri_rows"], fit["vt"].shape[0], np.median(np.abs(new - truth2) / truth2),
            spearmanr(new, truth2).correlation, M.get_auc_roc(new.astype(np.float32), y2)), flush=True)
