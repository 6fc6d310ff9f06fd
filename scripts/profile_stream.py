"""A few realtime-style host scoring calls (for ncu launch lists): python scripts/profile_stream.py BATCH CALLS PRECISION"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict
B, calls, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
D = 1728
m = get_model(argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision=prec)).eval()
m.load_state_dict(synth_state_dict(D, 100, 5, 0))
eng = m.engine()
x = np.random.default_rng(0).random((B, D), dtype=np.float32)
ts = []
for _ in range(calls):
    t0 = time.perf_counter(); eng.score_host(x, 0, 6, base=True, sap=True, nap=False); ts.append(time.perf_counter() - t0)
print("p50 us", sorted(ts)[len(ts) // 2] * 1e6)
