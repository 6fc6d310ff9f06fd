"""In-kernel time stamps of every tcgen05 GEMM of one train step (MMAD_TC_DEBUG=1, graphs off so each launch can be read back).
The third step is printed (the first two warm the caches)."""
import argparse, os, sys, types
os.environ["MMAD_NO_GRAPHS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
from icra2021_multimodal_ad_b200.optim import Adam
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = argparse.Namespace(input_size=1728, btl_size=100, n_layers=5, gpu_id=0, precision="f16x3")
model = get_model(cfg)
model.load_state_dict(synth_state_dict(1728, 100, 5, 0))
eng = types.SimpleNamespace(model=model, optimizer=Adam(model.parameters(), lr=1e-3), config=cfg)
x, _ = synth_windows(B, 1728, 1, anomaly_rate=0.0)
x = x.cuda()
for i in range(2):
    AutoEncoder.step(eng, (x, None))
torch.cuda.synchronize()
os.environ["MMAD_TC_DEBUG"] = "1"
sys.stderr.write("---- step 3 ----\n")
AutoEncoder.step(eng, (x, None))
torch.cuda.synchronize()
