"""Per-phase timing of the one-launch realtime kernel (csrc/stream.cu): run with MMAD_STREAM_DEBUG=1, the library prints
CTA 0's globaltimer stamps per phase (staging, then every layer step including its grid barrier) for each call.
python scripts/profile_stream.py"""
import argparse, os, sys, time
import numpy as np
os.environ.setdefault("MMAD_STREAM_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict

D = 1728
m = get_model(argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision="f16x3")).eval()
m.load_state_dict(synth_state_dict(D, 100, 5, 0))
eng = m.engine()
for b in (1, 4, 10, 16, 64):
    x = np.ascontiguousarray(np.random.default_rng(b).random((b, D), dtype=np.float32))
    for _ in range(20):
        eng.score_host(x, 0, 6)
    sys.stderr.flush()
    print(f"--- rows {b} (last calls above)", file=sys.stderr, flush=True)
