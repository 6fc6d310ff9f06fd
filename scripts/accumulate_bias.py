"""Direct look at the tensor core's fp32 accumulation: z = x W^T with ALL products positive (x, W > 0), one Linear layer,
K = 1728, against fp64.  A round-to-nearest FMA chain errs by a zero-mean random walk; an accumulator that is TRUNCATED
after every MMA instruction errs towards zero by ~0.5 ulp per instruction: a negative, systematic relative error that grows
with the number of instructions per output (f16x3: 3 per 16 k; f16f8: 1 fp16 per 16 k + 1 fp8 per 32 twin bytes).
python scripts/accumulate_bias.py"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model

D, btl, n = 1728, 100, 4096
g = torch.Generator().manual_seed(0)
x = torch.rand(n, D, generator=g) + 0.5
for prec in ("fp32", "f16x3", "f16f8", "f16"):
    m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=1, gpu_id=0, precision=prec)).eval()
    sd = m.state_dict()
    gw = torch.Generator().manual_seed(1)
    sd["encoder.net.0.layer.weight"] = (torch.rand(btl, D, generator=gw) + 0.5) / D
    sd["encoder.net.0.layer.bias"] = torch.zeros(btl)
    m.load_state_dict(sd)
    with torch.no_grad():
        z = m.encode(x.cuda()).double().cpu()
    ref = x.double() @ sd["encoder.net.0.layer.weight"].double().t()
    rel = ((z - ref) / ref).numpy()
    print(f"{prec:6s} relative error of z: mean {rel.mean():+.3e}  std {rel.std():.3e}  max|.| {np.abs(rel).max():.3e}   (2^-24 = 5.96e-08)")
