"""Does the hybrid NAP split point have room below 1e-5?  Quality on the golden all-layers case and rows / ms on bench.py's model
for tau in 1e-5 .. 1e-7 (f16x3).  python scripts/nap_tau_low.py"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from scipy.stats import spearmanr
import test_gpu_parity_r2 as T
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils import metric as M
from icra2021_multimodal_ad_b200.utils.synth import synth_windows
import bench

TAUS = (1e-5, 3e-6, 1e-6, 3e-7, 1e-7)
g, sd, xtr, xte, y, truth = T._nap_full_case()
D, btl, nl = g["D"], g["btl"], g["n_layers"]
eng = T._model(D, btl, nl, sd, "f16x3").engine()
for tau in TAUS:
    fit = eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False, factor="hybrid", tau=tau)
    new = eng.score(xte.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
    print("golden tau %.0e: triangular rows %d / %d  err %.3f rho %.4f auroc %.4f" % (
        tau, fit["tri_rows"], fit["vt"].shape[0], np.median(np.abs(new - truth) / truth), spearmanr(new, truth).correlation,
        M.get_auc_roc(new.astype(np.float32), y)), flush=True)
m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision="f16x3")).eval()
m.load_state_dict(bench.trained_state_dict(0))
eng = m.engine()
xs, _ = synth_windows(8192, D, 1234, anomaly_rate=0.0)
xf = xs.repeat(8, 1)
xf = (xf + 1e-3 * torch.randn(xf.shape, generator=torch.Generator().manual_seed(0))).clamp_(0, 1).cuda()
xb, _ = synth_windows(8192, D, 1236)
xb = xb.repeat(10, 1)[:75776].contiguous().cuda()
for tau in TAUS:
    fit = eng.nap_fit(xf, 0, nl + 1, distributed=False, factor="hybrid", tau=tau)
    for _ in range(2):
        eng.score(xb, 0, nl + 1, nap=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.score(xb, 0, nl + 1, nap=True)
    e1.record(); torch.cuda.synchronize()
    print("bench model tau %.0e: triangular rows %d / %d  %.2f ms/step" % (tau, fit["tri_rows"], fit["vt"].shape[0], e0.elapsed_time(e1) / 5), flush=True)
