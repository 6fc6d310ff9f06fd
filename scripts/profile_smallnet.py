"""Driver for profiling the fused tensor-core small-net kernel: `--stamps` makes the library print per-step time stamps
(MMAD_SNT_DEBUG=1); without it the script is the target of an ncu capture (-k regex:smallnet_tc)."""
import argparse, os, sys
if "--stamps" in sys.argv:
    os.environ["MMAD_SNT_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

for d in (64, 128):
    m = get_model(argparse.Namespace(input_size=d, btl_size=100, n_layers=5, gpu_id=0, precision="f16x3")).eval()
    m.load_state_dict(synth_state_dict(d, 100, 5, 0))
    x, _ = synth_windows(8192, d, 9)
    x = x.repeat(10, 1)[: 4 * 148 * 128].contiguous().cuda()
    eng = m.engine()
    for _ in range(2):
        eng.score(x, 0, 6)
    torch.cuda.synchronize()
