"""Evidence for the training parity protocol: perturbing the weights by 6e-8 relative moves the
REFERENCE ALGORITHM's own gradients (oracle, fp32 CPU) by either ~2e-6 (no LeakyReLU branch flip) or ~2e-3 of
the tensor max on every upstream tensor (one flip).  CPU only; run from the repo root."""
import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
from oracle import rapp_oracle as RO
torch.set_num_threads(8)
D, btl, nl, seed, B = 1728, 100, 5, 51, 256
sd = synth_state_dict(D, btl, nl, seed)
x, _ = synth_windows(B, D, seed + 100, anomaly_rate=0.0)
l0, g0, _ = RO.train_forward_backward(x, dict(sd))
for trial in range(3):
    sd2 = {k: (v * (1 + 6e-8 * torch.randn(v.shape, generator=torch.Generator().manual_seed(trial))) if v.dtype.is_floating_point and v.dim()==2 else v) for k, v in sd.items()}
    l1, g1, _ = RO.train_forward_backward(x, sd2)
    import numpy as np
    es = []
    for k in ["encoder.net.0.bn.bias","encoder.net.0.layer.weight","encoder.net.3.layer.weight","decoder.net.1.layer.weight","decoder.net.3.layer.weight"]:
        s = g0[k].abs().max()
        e = ((g1[k]-g0[k]).abs()/s)
        es.append("%s max %.1e q98 %.1e med %.1e" % (k.replace(".net.","").replace("layer.",""), e.max(), torch.quantile(e.flatten()[:1000000],0.98), e.median()))
    print(trial, " | ".join(es))
