"""Host-call latency (mmad_score_host, base + SAP) for small batches: python scripts/time_small_batches.py [rows ...]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict

D, BTL, NL = 1728, 100, 5
m = get_model(argparse.Namespace(input_size=D, btl_size=BTL, n_layers=NL, gpu_id=0, precision="f16x3")).eval()
m.load_state_dict(synth_state_dict(D, BTL, NL, 0))
eng = m.engine()
for b in [int(a) for a in sys.argv[1:]] or [10, 17, 32, 64, 128, 256, 512]:
    x = np.ascontiguousarray(np.random.default_rng(b).random((b, D), dtype=np.float32))
    for _ in range(40):
        eng.score_host(x, 0, NL + 1, base=True, sap=True, nap=False)
    ts = []
    for _ in range(300):
        t0 = time.perf_counter()
        eng.score_host(x, 0, NL + 1, base=True, sap=True, nap=False)
        ts.append(time.perf_counter() - t0)
    ts = np.sort(np.asarray(ts)) * 1e6
    print(f"rows {b:4d}: p50 {ts[len(ts) // 2]:7.1f} us  p99 {ts[int(len(ts) * 0.99)]:7.1f} us", flush=True)
