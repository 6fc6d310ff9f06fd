"""Run a few fused train steps (for ncu launch lists): python scripts/profile_train.py BATCH STEPS PRECISION"""
import argparse, os, sys, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
from icra2021_multimodal_ad_b200.optim import Adam
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

B, steps, prec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
D = 1728
cfg = argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision=prec)
m = get_model(cfg)
m.load_state_dict(synth_state_dict(D, 100, 5, 0))
eng = types.SimpleNamespace(model=m, optimizer=Adam(m.parameters(), lr=1e-3), config=cfg)
x, _ = synth_windows(B, D, 1, anomaly_rate=0.0)
x = x.cuda()
for _ in range(steps):
    AutoEncoder.step(eng, (x, None))
torch.cuda.synchronize()
print("done")
