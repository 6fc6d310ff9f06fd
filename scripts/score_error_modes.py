"""Per-sample score error of every precision mode against the CPU oracle (fp32): python scripts/score_error_modes.py [D] [N]."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
from oracle import rapp_oracle as RO

D = int(sys.argv[1]) if len(sys.argv) > 1 else 1728
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
sd = synth_state_dict(D, 100, 5, 0)
x, _ = synth_windows(N, D, 1236)
ref = RO.get_diffs(x, sd, batch_size=256)
sap_o, base_o = RO.sap_score(ref).astype(np.float64), RO.recon_score(ref[0]).astype(np.float64)
for prec in ("fp32", "f16x3", "f16f8", "f16"):
    m = get_model(argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision=prec)).eval()
    m.load_state_dict(sd)
    o = m.engine().score(x.cuda(), 0, 6)
    sap, base = o["sap"].cpu().numpy().astype(np.float64), o["base"].cpu().numpy().astype(np.float64)
    es, eb = np.abs(sap - sap_o) / sap_o, np.abs(base - base_o) / base_o
    print(f"D={D} N={N} {prec:6s} vs oracle fp32: SAP max {es.max():.2e} med {np.median(es):.2e} | base max {eb.max():.2e} med {np.median(eb):.2e}")
