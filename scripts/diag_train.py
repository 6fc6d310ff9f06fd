"""Diagnostic: per-parameter gradient error of the fused train step vs the oracle in fp32 and fp64."""
import argparse, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
from icra2021_multimodal_ad_b200.model_builder import get_model
from oracle import rapp_oracle as RO

D, btl, nl, seed, B = 1728, 100, 5, 51, 256
prec, xseed = "fp32", None
if len(sys.argv) > 1:
    D, B, seed, xseed, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
sd = synth_state_dict(D, btl, nl, seed)
x, _ = synth_windows(B, D, seed + 100 if xseed is None else xseed, anomaly_rate=0.0)
l32, g32, _ = RO.train_forward_backward(x, dict(sd))
sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
l64, g64, _ = RO.train_forward_backward(x.double(), sd64)
m = get_model(argparse.Namespace(input_size=D, btl_size=btl, n_layers=nl, gpu_id=0, precision=prec))
m.load_state_dict(sd); m.train()
loss = m.get_loss_value(x.cuda(), x.cuda()); loss.backward()
print("loss", float(loss), l32, l64)
for k, p in m.named_parameters():
    t = g64[k]; s = t.abs().max().item()
    e_mine = (p.grad.cpu().double() - t).abs().max().item() / s
    e_ref = (g32[k].double() - t).abs().max().item() / s
    print("%-32s mine %.2e  oracle32 %.2e" % (k, e_mine, e_ref))
