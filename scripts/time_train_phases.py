"""Where does a train step go?  python scripts/time_train_phases.py BATCH PRECISION"""
import argparse, os, sys, time, types
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.model_builder import get_model
from icra2021_multimodal_ad_b200.optim import Adam
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

B, prec = int(sys.argv[1]), sys.argv[2]
D = 1728
cfg = argparse.Namespace(input_size=D, btl_size=100, n_layers=5, gpu_id=0, precision=prec)
m = get_model(cfg); m.load_state_dict(synth_state_dict(D, 100, 5, 0)); m.train()
opt = Adam(m.parameters(), lr=1e-3)
x, _ = synth_windows(B, D, 1, anomaly_rate=0.0); x = x.cuda()

def phase(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, th / n * 1e6

def f_loss():
    opt.zero_grad()
    return m.get_loss_value(x, x)
def f_lossbwd():
    l = f_loss(); l.backward(); return l
def f_full():
    l = f_lossbwd(); opt.step(); return float(l.detach())
f_lossbwd()
def f_adam(): opt.step()
for name, fn in (("loss(fwd+bwd kernels)", f_loss), ("loss+autograd backward", f_lossbwd), ("adam only", f_adam), ("full step + float(loss)", f_full)):
    dev, host = phase(fn)
    print("%-28s device %8.1f us   host-issue %8.1f us" % (name, dev, host))
