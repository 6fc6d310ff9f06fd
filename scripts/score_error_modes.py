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
STEPS = int(os.environ.get("TRAIN_STEPS", "0"))     # oracle Adam steps before scoring (a trained model is the hard case)
if STEPS:
    xtr, _ = synth_windows(256 * 8, D, 7, anomaly_rate=0.0)
    opt = {}
    for i in range(STEPS):
        RO.train_step(xtr[(i % 8) * 256:(i % 8 + 1) * 256], sd, opt)
x, _ = synth_windows(N, D, 1236)
ref = RO.get_diffs(x, sd, batch_size=256)
sap_o, base_o = RO.sap_score(ref).astype(np.float64), RO.recon_score(ref[0]).astype(np.float64)
sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
with torch.no_grad():
    enc64 = RO.module_layers(sd64, "encoder")
    xd = x.double(); xt = RO.ae_forward_eval(xd, sd64); d64 = [xt - xd]
    for L in enc64:
        xd = RO.fc_layer_eval(xd, L); xt = RO.fc_layer_eval(xt, L); d64.append(xt - xd)
sap64 = torch.cat(d64, 1).pow(2).mean(1).numpy(); base64 = d64[0].pow(2).mean(1).numpy()
print(f"oracle fp32 vs fp64: SAP max {np.abs(sap_o - sap64).max() / 1:.2e} rel {(np.abs(sap_o - sap64) / sap64).max():.2e} | base rel {(np.abs(base_o - base64) / base64).max():.2e}")
for prec in ("fp32", "f16x3", "f16f8", "f16"):
    m = get_model(argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision=prec)).eval()
    m.load_state_dict(sd)
    o = m.engine().score(x.cuda(), 0, 6)
    sap, base = o["sap"].cpu().numpy().astype(np.float64), o["base"].cpu().numpy().astype(np.float64)
    es, eb = np.abs(sap - sap_o) / sap_o, np.abs(base - base_o) / base_o
    e64 = np.abs(sap - sap64) / sap64
    print(f"D={D} N={N} steps={STEPS} {prec:6s} vs oracle fp32: SAP max {es.max():.2e} med {np.median(es):.2e} | base max {eb.max():.2e} med {np.median(eb):.2e} | vs fp64: SAP max {e64.max():.2e} med {np.median(e64):.2e}")
