"""Host-buffer scoring throughput (mmad_score_host: pinned host x -> scores on the host): python scripts/time_e2e.py [PRECISION] [ROWS] [CALLS]"""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

prec = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 75776
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 8
D = 1728
m = get_model(argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision=prec)).eval()
m.load_state_dict(synth_state_dict(D, 100, 5, 0))
eng = m.engine()
xtr, _ = synth_windows(8192, D, 1234, anomaly_rate=0.0)
eng.nap_fit(xtr.cuda(), 0, 6, distributed=False)
xs, _ = synth_windows(8192, D, 1236)
x = xs.repeat((rows + 8191) // 8192, 1)[:rows].contiguous().pin_memory().numpy()
for _ in range(2):
    eng.score_host(x, 0, 6, base=True, sap=True, nap=True)
torch.cuda.synchronize()
ts = []
for _ in range(calls):
    t0 = time.perf_counter()
    eng.score_host(x, 0, 6, base=True, sap=True, nap=True)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts)
print(f"{prec} rows={rows} tail={os.environ.get('MMAD_HOST_TAIL', 'default')}: median {np.median(ts) * 1e3:.2f} ms/call = {rows / np.median(ts) / 1e6:.2f} M windows/s (best {rows / ts.min() / 1e6:.2f})")
