"""CPU emulation of the tensor-core operand splits for the scoring chain (which mode can hold the 1e-4 score bar).

  f16x3 : hi*hi + hi*lo + lo*hi, fp16 twins (the shipped fp32-parity mode)
  f16f8 : hi*hi in fp16 + [lo8 | a8] . [Wh8 ; Wl8] in fp8 e4m3 (one f16 pass + one double-length fp8 pass)
  f16x2 : (hi+lo)*Wh  (A exact, W rounded to fp16)
  f16   : hi*hi

Products are exact in the tensor core and accumulation is fp32; here products/accumulation run in fp64 and the
result is rounded to fp32, so only the operand quantisation is modelled.
"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
from oracle import rapp_oracle as RO

F8 = torch.float8_e4m3fn


def q8(x):
    return x.clamp(-448, 448).to(F8).to(torch.float64)


def q16(x):
    return x.to(torch.float32).clamp(-65504, 65504).to(torch.float16).to(torch.float64)


def pow2_scale(maxabs, top):
    """power of two s with maxabs*s in [top/2, top)"""
    return 2.0 ** np.floor(np.log2(top / max(maxabs, 1e-30)))


def mm(a, W, mode, wtop=2.0 ** 14):
    """a [B,K] fp32, W [N,K] fp32 -> a W^T fp32 under the given operand split."""
    a64, W64 = a.double(), W.double()
    if mode == "fp32":
        return (a @ W.t())
    if mode in ("f16x3", "f16x2", "f16"):
        ah = q16(a64); al = q16(a64 - ah)
        Ws = W64 * 256.0
        wh = q16(Ws); wl = q16(Ws - wh)
        if mode == "f16":
            acc = ah @ wh.t()
        elif mode == "f16x2":
            acc = (ah + al) @ wh.t()
        else:
            acc = ah @ wh.t() + ah @ wl.t() + al @ wh.t()
        return (acc / 256.0).float()
    if mode == "f16f8":
        sW = pow2_scale(float(W.abs().max()), wtop)
        ah = q16(a64); al = a64 - ah
        Ws = W64 * sW
        wh = q16(Ws); wl = Ws - wh
        a2 = 2.0 ** 11
        al8 = q8(al * a2); wh8 = q8(wh / a2)
        a8 = q8(a64); wl8 = q8(wl)
        acc = ah @ wh.t() + al8 @ wh8.t() + a8 @ wl8.t()
        return (acc / sW).float()
    raise ValueError(mode)


def layer(x, L, mode):
    y = mm(x, L["W"], mode) + L["b"]
    if "gamma" in L:
        y = torch.nn.functional.leaky_relu(y, 0.2)
        sc = L["gamma"] / torch.sqrt(L["var"] + 1e-5)
        y = y * sc + (L["beta"] - L["mean"] * sc)
    return y


def diffs(x, sd, mode):
    enc, dec = RO.module_layers(sd, "encoder"), RO.module_layers(sd, "decoder")
    h = x; H = []
    for L in enc:
        h = layer(h, L, mode); H.append(h)
    for L in dec:
        h = layer(h, L, mode)
    out = [h - x]
    for L, ref in zip(enc, H):
        h = layer(h, L, mode); out.append(h - ref)
    return out


def main():
    D = int(os.environ.get("D", 1728)); steps = int(os.environ.get("TRAIN_STEPS", 60)); n = int(os.environ.get("N", 1024))
    torch.manual_seed(0)
    sd = synth_state_dict(D, 100, 5, 0)
    if steps:
        xtr, _ = synth_windows(256 * 8, D, 7, anomaly_rate=0.0)
        opt = {}
        for i in range(steps):
            xb = xtr[(i % 8) * 256:(i % 8 + 1) * 256]
            RO.train_step(xb, sd, opt)
    x, _ = synth_windows(n, D, 1236)
    ref = diffs(x.double(), {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, "fp32")
    ref_sap = torch.cat(ref, 1).pow(2).mean(1); ref_base = ref[0].pow(2).mean(1)
    for mode in ("fp32", "f16x3", "f16f8", "f16x2", "f16"):
        d = diffs(x, sd, mode)
        sap = torch.cat(d, 1).double().pow(2).mean(1); base = d[0].double().pow(2).mean(1)
        es = ((sap - ref_sap).abs() / ref_sap); eb = ((base - ref_base).abs() / ref_base)
        ed = max(float((a.double() - b).abs().max() / b.abs().max()) for a, b in zip(d, ref))
        print(f"{mode:6s} SAP rel err max {float(es.max()):.2e} med {float(es.median()):.2e} | base max {float(eb.max()):.2e} | diffs/max {ed:.2e}")


if __name__ == "__main__":
    main()
