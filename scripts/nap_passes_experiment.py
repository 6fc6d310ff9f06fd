"""What does the two-pass NAP rotation (mmad_set_option "nap_passes" = 2: whitening rows rounded to fp16) cost in accuracy?
Runs the two NAP parity protocols of tests/test_gpu_parity_r2.py with 3 and 2 passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from scipy.stats import spearmanr
import test_gpu_parity_r2 as T
from icra2021_multimodal_ad_b200.utils import metric as M

g, sd, xtr, xte, y, truth = T._nap_full_case()
D, btl, nl = g["D"], g["btl"], g["n_layers"]
ref = g["nap"]["score"].numpy().astype(np.float64)
for passes in (3, 2, 4):
    eng = T._model(D, btl, nl, sd, "f16x3").engine()
    eng.set_option("nap_passes", passes)
    eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False)
    new = eng.score(xte.cuda(), 0, nl + 1, base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
    ok = np.isfinite(new) & np.isfinite(ref)
    print("all layers, passes %d: err_new %.3g err_ref %.3g rho_new %.4f rho_ref %.4f auroc new %.4f ref %.4f" % (
        passes, np.median(np.abs(new[ok] - truth[ok]) / truth[ok]), np.median(np.abs(ref[ok] - truth[ok]) / truth[ok]),
        spearmanr(new[ok], truth[ok]).correlation, spearmanr(ref[ok], truth[ok]).correlation,
        M.get_auc_roc(new.astype(np.float32), y), g["nap"]["metrics"][0]), flush=True)
for name, sel in (("score_D1728.pt", (0, 1)), ("score_D1728.pt", (5, 6)), ("score_D64.pt", (0, 1))):
    g, sd, xtr, xte, y, truth = T._nap_case(name, sel)
    D, btl, nl = g["D"], g["btl"], g["n_layers"]
    for passes in (3, 2, 4):
        eng = T._model(D, btl, nl, sd, "f16x3").engine()
        eng.set_option("nap_passes", passes)
        eng.nap_fit(xtr.cuda(), sel[0], sel[1], distributed=False)
        s = eng.score(xte.cuda(), sel[0], sel[1], base=False, sap=False, nap=True)["nap"].cpu().numpy().astype(np.float64)
        err = np.abs(s - truth) / truth
        print(f"{name} {sel} passes {passes}: max {err.max():.2e} median {np.median(err):.2e}", flush=True)
