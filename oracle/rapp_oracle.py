"""CPU restatement of the reference autoencoder + RaPP scorer.  TEST INFRASTRUCTURE.

Every function names the reference file:line it follows (paths relative to the
reference checkout).  Arithmetic is torch-CPU fp32 functional ops, i.e. the same
library calls the reference's ``nn.Module`` objects make, written without any
of the reference's classes.  Pinned against ``tests/golden/*.pt`` (see
``oracle/__init__.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2      # modules/activation.py:37-38  nn.LeakyReLU(.2)
BN_EPS = 1e-5          # torch BatchNorm1d default, layers/fc_layer.py:30
BN_MOMENTUM = 0.1      # torch BatchNorm1d default


# ----------------------------------------------------------------------------
# shapes
# ----------------------------------------------------------------------------
def hidden_layer_sizes(start_size: int, end_size: int, n_hidden_layers: int) -> List[int]:
    """utils/common_utils.py:22-31 -- linear interpolation, truncating int()."""
    diff = (start_size - end_size) / (n_hidden_layers + 1)
    return [int(start_size - diff * (i + 1)) for i in range(n_hidden_layers)]


def encoder_widths(input_size: int, btl_size: int, n_layers: int) -> List[int]:
    """model_builder.py:21-28 -- [D, h1, .., h_{n-1}, btl]."""
    return [input_size] + hidden_layer_sizes(input_size, btl_size, n_layers - 1) + [btl_size]


def decoder_widths(input_size: int, btl_size: int, n_layers: int) -> List[int]:
    """model_builder.py:30-37 -- [btl, .., D]; computed separately (not mirrored)."""
    return [btl_size] + hidden_layer_sizes(btl_size, input_size, n_layers - 1) + [input_size]


# ----------------------------------------------------------------------------
# state-dict access (key names: SURVEY.md section 5, [probe] listing)
# ----------------------------------------------------------------------------
def module_layers(sd: Dict[str, torch.Tensor], prefix: str) -> List[dict]:
    """Collect ``{prefix}.net.{i}.layer.{weight,bias}`` / ``.bn.*`` into a list."""
    layers = []
    i = 0
    while f"{prefix}.net.{i}.layer.weight" in sd:
        ent = {"W": sd[f"{prefix}.net.{i}.layer.weight"], "b": sd[f"{prefix}.net.{i}.layer.bias"]}
        if f"{prefix}.net.{i}.bn.weight" in sd:
            ent.update(gamma=sd[f"{prefix}.net.{i}.bn.weight"], beta=sd[f"{prefix}.net.{i}.bn.bias"],
                       mean=sd[f"{prefix}.net.{i}.bn.running_mean"], var=sd[f"{prefix}.net.{i}.bn.running_var"])
        layers.append(ent)
        i += 1
    return layers


# ----------------------------------------------------------------------------
# forward (eval mode)
# ----------------------------------------------------------------------------
def fc_layer_eval(x: torch.Tensor, L: dict) -> torch.Tensor:
    """layers/fc_layer.py:37-48 in eval mode: Linear -> LeakyReLU(.2) -> BatchNorm
    (running stats) for hidden layers, bare Linear for the last one
    (modules/fc_module.py:42-56).  >2-D input is flattened for BN (41-43)."""
    y = F.linear(x, L["W"], L["b"])
    if "gamma" in L:
        y = F.leaky_relu(y, LRELU_SLOPE)
        shp = y.shape
        y = F.batch_norm(y.reshape(-1, shp[-1]), L["mean"], L["var"], L["gamma"], L["beta"],
                         training=False, momentum=BN_MOMENTUM, eps=BN_EPS).reshape(shp)
    return y


def module_forward_eval(x: torch.Tensor, layers: Sequence[dict]) -> torch.Tensor:
    """modules/fc_module.py:59-61 (distribution=None passthrough of the decorator)."""
    for L in layers:
        x = fc_layer_eval(x, L)
    return x


def ae_forward_eval(x: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """models/auto_encoder.py:36-50."""
    z = module_forward_eval(x, module_layers(sd, "encoder")).view(x.size(0), -1)
    return module_forward_eval(z, module_layers(sd, "decoder")).view(x.size(0), -1)


def recon_loss_sum(x: torch.Tensor, sd) -> float:
    """models/auto_encoder.py:52-55,79-91 + modules/loss.py:31-32 -- MSE, reduction='sum'."""
    return float(F.mse_loss(ae_forward_eval(x, sd), x, reduction="sum"))


# ----------------------------------------------------------------------------
# RaPP diffs and scores
# ----------------------------------------------------------------------------
def get_diffs(x, sd, batch_size: int = 698) -> List[np.ndarray]:
    """reconstruction_aggregation.py:6-37.  d_0 = x_hat - x; d_l = enc_l(x_hat) - enc_l(x)."""
    if isinstance(x, np.ndarray):
        x = torch.tensor(x)
    enc = module_layers(sd, "encoder")
    per_batch = []
    with torch.no_grad():
        for xb in x.split(batch_size):
            xb = xb.float()
            xt = ae_forward_eval(xb, sd)
            diffs = [xt - xb]
            for L in enc:
                xb = fc_layer_eval(xb, L)
                xt = fc_layer_eval(xt, L)
                diffs.append(xt - xb)
            per_batch.append(diffs)
    return [torch.cat(s, dim=0).numpy() for s in zip(*per_batch)]


def clamp_layer_range(n_diffs: int, start: int, end: Optional[int]) -> Tuple[int, int]:
    """utils/metric.py:155-162 and 195-202 (identical in SAP and NAP)."""
    if end is None:
        end = n_diffs + 1
    if start > n_diffs - 1:
        start = n_diffs - 1
    if end - start < 1:
        end = start + 1
    return start, end


def concat_diffs(diffs: Sequence[np.ndarray], start: int = 0, end: Optional[int] = None) -> np.ndarray:
    """utils/metric.py:166-167,204-209 -- python slice then concat on the last dim."""
    start, end = clamp_layer_range(len(diffs), start, end)
    return np.concatenate([np.asarray(d) for d in diffs[start:end]], axis=-1)


def recon_score(d0: np.ndarray) -> np.ndarray:
    """utils/metric.py:132-133 -- base score = mean_j d_0^2."""
    return (d0 ** 2).mean(axis=1)


def sap_score(diffs: Sequence[np.ndarray], start: int = 0, end: Optional[int] = None) -> np.ndarray:
    """utils/metric.py:145-171 -- SAP = mean over the concatenated dims of d^2."""
    return (concat_diffs(diffs, start, end) ** 2).mean(axis=1)


class NapFit:
    """utils/normalize.py:47-70 (Rotater.fit) + 20-34 (Standardizer.fit), fp32 like the reference."""

    def __init__(self, train_concat: np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(train_concat)).float()
        self.mu = x.mean(dim=0)
        xc = x - self.mu
        _, self.s, self.v = xc.svd()
        rot = torch.matmul(xc, self.v)
        self.mu2 = rot.mean(dim=0)
        r = rot - self.mu2
        self.var = torch.from_numpy(np.cov(r.numpy().T)).reshape(r.shape[1], -1).diagonal().float() \
            if r.shape[1] > 1 else torch.from_numpy(np.atleast_1d(np.cov(r.numpy().T))).float()
        self.n_train = x.shape[0]

    def score(self, concat: np.ndarray) -> np.ndarray:
        """utils/normalize.py:72-103,36-45 + utils/metric.py:220-222."""
        x = torch.from_numpy(np.ascontiguousarray(concat)).float()
        rot = torch.matmul(x - self.mu, self.v)
        z = ((rot - self.mu2) / self.var ** .5).numpy()
        return (abs(z) ** 2).mean(axis=1)


def nap_score(train_diffs, test_diffs, start: int = 0, end: Optional[int] = None) -> np.ndarray:
    """utils/metric.py:183-222."""
    fit = NapFit(concat_diffs(train_diffs, start, end))
    return fit.score(concat_diffs(test_diffs, start, end))


def nap_score_fp64(train_concat: np.ndarray, test_concat: np.ndarray) -> np.ndarray:
    """fp64 closed form of the same quantity:
    NAP(x) = (N-1)/K * sum_j ((x-mu).v_j - mu2_j)^2 / lambda_j with G = Xc^T Xc = V L V^T.
    Used as 'truth' for the ill-conditioned all-layers selection."""
    xt = np.asarray(train_concat, dtype=np.float64)
    mu = xt.mean(axis=0)
    xc = xt - mu
    _, s, vt = np.linalg.svd(xc, full_matrices=False)
    rot_tr = xc @ vt.T
    mu2 = rot_tr.mean(axis=0)
    var = ((rot_tr - mu2) ** 2).sum(axis=0) / (xt.shape[0] - 1)
    rot = (np.asarray(test_concat, dtype=np.float64) - mu) @ vt.T
    return (((rot - mu2) ** 2) / var).mean(axis=1)


# ----------------------------------------------------------------------------
# VIB reparameterisation
# ----------------------------------------------------------------------------
def vib_normal(output: torch.Tensor, eps: Optional[torch.Tensor], k: int = 1,
               stochastic: bool = True) -> Dict[str, torch.Tensor]:
    """decorators/variational_info_bottleneck.py:19-42 with the noise made an input
    (the reference draws ``torch.randn_like``)."""
    if k < 1:
        raise ValueError("k should be >= 1")
    mu, logvar = output.split(output.size(-1) // 2, dim=-1)
    sigma = (logvar * .5).exp()
    if stochastic:
        z = eps.mul(sigma.unsqueeze(0).expand(k, *sigma.size())) + mu
    else:
        z = mu.unsqueeze(0).expand(k, *mu.size())
    return {"z": z, "mu": mu, "logvar": logvar}


# ----------------------------------------------------------------------------
# training step: manual forward/backward + Adam (no autograd)
# ----------------------------------------------------------------------------
def _bn_train_fwd(a: torch.Tensor, gamma, beta):
    """torch BatchNorm1d training forward: biased batch variance for normalisation."""
    mean = a.mean(dim=0)
    var_b = a.var(dim=0, unbiased=False)
    inv = torch.rsqrt(var_b + BN_EPS)
    xhat = (a - mean) * inv
    return xhat * gamma + beta, (xhat, inv, mean, var_b)


def train_forward_backward(x: torch.Tensor, sd: Dict[str, torch.Tensor]):
    """models/auto_encoder.py:57-77 minus the optimizer: train-mode forward (BN batch
    statistics), loss = sum (x_hat-x)^2, gradients of every parameter, and the BN
    running-stat updates (momentum .1, unbiased running var, num_batches_tracked+1).
    Returns (loss, grads{key}, new_buffers{key})."""
    B = x.shape[0]
    saved = []
    h = x
    names = []
    for prefix in ("encoder", "decoder"):
        for i, L in enumerate(module_layers(sd, prefix)):
            names.append((prefix, i))
            pre = F.linear(h, L["W"], L["b"])
            if "gamma" in L:
                a = F.leaky_relu(pre, LRELU_SLOPE)
                out, (xhat, inv, mean, var_b) = _bn_train_fwd(a, L["gamma"], L["beta"])
                saved.append(dict(inp=h, pre=pre, xhat=xhat, inv=inv, mean=mean, var_b=var_b, L=L))
            else:
                out = pre
                saved.append(dict(inp=h, pre=pre, L=L))
            h = out
    diff = h - x
    loss = float((diff * diff).sum())
    g = 2.0 * diff
    grads: Dict[str, torch.Tensor] = {}
    bufs: Dict[str, torch.Tensor] = {}
    for (prefix, i), S in reversed(list(zip(names, saved))):
        L = S["L"]
        key = f"{prefix}.net.{i}"
        if "gamma" in L:
            xhat, inv = S["xhat"], S["inv"]
            grads[key + ".bn.weight"] = (g * xhat).sum(dim=0)
            grads[key + ".bn.bias"] = g.sum(dim=0)
            gx = g * L["gamma"]
            g = inv / B * (B * gx - gx.sum(dim=0) - xhat * (gx * xhat).sum(dim=0))
            g = torch.where(S["pre"] > 0, g, g * LRELU_SLOPE)
            bufs[key + ".bn.running_mean"] = (1 - BN_MOMENTUM) * L["mean"] + BN_MOMENTUM * S["mean"]
            bufs[key + ".bn.running_var"] = (1 - BN_MOMENTUM) * L["var"] + BN_MOMENTUM * S["var_b"] * (B / (B - 1))
            bufs[key + ".bn.num_batches_tracked"] = sd[key + ".bn.num_batches_tracked"] + 1
        grads[key + ".layer.weight"] = g.t() @ S["inp"]
        grads[key + ".layer.bias"] = g.sum(dim=0)
        g = g @ L["W"]
    return loss, grads, bufs


def adam_update(p, g, m, v, step: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (novelty_detection.py:90; weight_decay 0, amsgrad off), single-tensor
    formulation: m.lerp_(g, 1-b1); v = b2 v + (1-b2) g^2;
    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""
    m = m + (g - m) * (1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2_sqrt = math.sqrt(1 - b2 ** step)
    denom = v.sqrt() / bc2_sqrt + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def train_step(x: torch.Tensor, sd: Dict[str, torch.Tensor], opt: Dict[str, dict], lr=1e-3):
    """One full models/auto_encoder.py:57-77 step.  ``opt`` maps parameter key ->
    {'m','v','step'} (created on first use).  Mutates ``sd``/``opt``; returns the loss."""
    loss, grads, bufs = train_forward_backward(x, sd)
    for k, g in grads.items():
        st = opt.setdefault(k, {"m": torch.zeros_like(sd[k]), "v": torch.zeros_like(sd[k]), "step": 0})
        st["step"] += 1
        sd[k], st["m"], st["v"] = adam_update(sd[k], g, st["m"], st["v"], st["step"], lr=lr)
    sd.update(bufs)
    return loss


# ----------------------------------------------------------------------------
# VIB autoencoder training (BASELINE configs[3]) -- KL term PARITY-UNPINNED
# ----------------------------------------------------------------------------
def vib_train_forward_backward(x: torch.Tensor, sd: Dict[str, torch.Tensor], eps: torch.Tensor, beta_kl: float):
    """Train-mode step of an autoencoder whose encoder output is split (mu | logvar) and
    reparameterised exactly as decorators/variational_info_bottleneck.py:19-27 does (k = 1, noise given),
    loss = sum (x_hat-x)^2 (modules/loss.py:31-32) + beta_kl * KL, KL = -1/2 sum(1+logvar-mu^2-exp(logvar)).
    The reference defines no KL term and no VIB model: the reparameterisation is pinned by
    tests/golden/vib_D64.pt, the KL by nothing -- this torch-autograd fp32 statement is the checker.
    Returns (loss, grads{key}, new_buffers{key})."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    bufs: Dict[str, torch.Tensor] = {}

    def run(h, prefix):
        i = 0
        while f"{prefix}.net.{i}.layer.weight" in sd:
            key = f"{prefix}.net.{i}"
            h = F.linear(h, params[key + ".layer.weight"], params[key + ".layer.bias"])
            if key + ".bn.weight" in sd:
                h = F.leaky_relu(h, LRELU_SLOPE)
                rm, rv = sd[key + ".bn.running_mean"].clone(), sd[key + ".bn.running_var"].clone()
                h = F.batch_norm(h, rm, rv, params[key + ".bn.weight"], params[key + ".bn.bias"], training=True,
                                 momentum=BN_MOMENTUM, eps=BN_EPS)
                bufs[key + ".bn.running_mean"], bufs[key + ".bn.running_var"] = rm, rv
                bufs[key + ".bn.num_batches_tracked"] = sd[key + ".bn.num_batches_tracked"] + 1
            i += 1
        return h

    out = run(x, "encoder")
    r = vib_normal(out, eps.unsqueeze(0), k=1)
    xhat = run(r["z"][0], "decoder")
    kl = -0.5 * torch.sum(1 + r["logvar"] - r["mu"] ** 2 - r["logvar"].exp())
    loss = F.mse_loss(xhat, x, reduction="sum") + beta_kl * kl
    loss.backward()
    return float(loss), {k: p.grad for k, p in params.items()}, bufs
