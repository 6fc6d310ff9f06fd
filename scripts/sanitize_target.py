"""Small end-to-end target for compute-sanitizer (one tool per gpurun call): scoring in every mode incl. the CTA-pair kernel,
NAP fit + score, the one-launch realtime kernel, the fused small-net kernel, device metrics, one fused train step + Adam.
compute-sanitizer --tool memcheck python scripts/sanitize_target.py"""
import argparse, os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
from icra2021_multimodal_ad_b200.optim import Adam
from icra2021_multimodal_ad_b200.utils import metric as M
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

for D, btl, nl, n in ((300, 17, 2, 2048 + 77), (64, 100, 5, 500)):
    sd = synth_state_dict(D, btl, nl, 1)
    x, y = synth_windows(n, D, 2)
    xtr, _ = synth_windows(1024, D, 3, anomaly_rate=0.0)
    for prec in ("fp32", "f16x3", "f16f8"):
        m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=prec)).eval()
        m.load_state_dict(sd)
        eng = m.engine()
        eng.nap_fit(xtr.cuda(), 0, nl + 1, distributed=False)
        o = eng.score(x.cuda(), 0, nl + 1, nap=True, diffs=True)           # pair kernel (>= 2048 rows) / single-CTA kernel
        eng.score(x[:300].cuda(), 0, nl + 1)
        h = eng.score_host(x[:10].numpy(), 0, nl + 1)                       # one-launch realtime kernel
        eng.score_host(x[:1].numpy(), 0, nl + 1)
        eng.score_host(x[:40].numpy(), 0, nl + 1)                           # graph-replay path
        eng.score_host(x.numpy(), 0, nl + 1, nap=True)                      # pipelined bulk path
        print(D, prec, float(o["sap"].mean()), float(h["sap"].mean()), M.get_auc_roc(o["nap"], y.cuda()), flush=True)
    cfg = argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision="f16x3")
    m = get_model(cfg)
    m.load_state_dict(sd)
    e = types.SimpleNamespace(model=m, optimizer=Adam(m.parameters(), lr=1e-3), config=cfg)
    for _ in range(2):
        loss, = AutoEncoder.step(e, (x[:96], None))
    print(D, "train", loss, flush=True)
torch.cuda.synchronize()
print("sanitize target done")
