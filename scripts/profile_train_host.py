"""Host-side cost of a train step (cProfile over 300 steps): the step ends with a device->host read of the loss, so the host
time in front of the next launch is GPU idle time."""
import argparse, cProfile, os, pstats, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
from icra2021_multimodal_ad_b200.optim import Adam
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

cfg = argparse.Namespace(input_size=1728, btl_size=100, n_layers=5, gpu_id=0, precision="f16x3")
model = get_model(cfg)
model.load_state_dict(synth_state_dict(1728, 100, 5, 0))
eng = types.SimpleNamespace(model=model, optimizer=Adam(model.parameters(), lr=1e-3), config=cfg)
x, _ = synth_windows(256, 1728, 1, anomaly_rate=0.0)
x = x.cuda()
for _ in range(20):
    AutoEncoder.step(eng, (x, None))
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    AutoEncoder.step(eng, (x, None))
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
