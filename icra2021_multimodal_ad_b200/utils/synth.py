"""Seeded synthetic weights / windows shared by tests, golden generation and bench.

The reference's dataset is unpublished (README.md:15), so every measurement and
parity test uses synthetic concatenated multimodal vectors with the real column
layout ``[hand-RGB 0:1024 | depth 1024:1536 | force-torque 1536:1600 | mic 1600:1728]``
(utils/data_loaders.py:226,404) in the min-max-normalised range [0,1]
(utils/data_loaders.py:448-457).  Everything is generated on the CPU with
``torch.Generator`` so the same tensors exist on the build container (where the
reference runs) and on the GPU box (where it does not).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Tuple

import torch


def hidden_layer_sizes(start_size: int, end_size: int, n_hidden_layers: int) -> List[int]:
    diff = (start_size - end_size) / (n_hidden_layers + 1)
    return [int(start_size - diff * (i + 1)) for i in range(n_hidden_layers)]


def ae_widths(input_size: int, btl_size: int = 100, n_layers: int = 5) -> Tuple[List[int], List[int]]:
    enc = [input_size] + hidden_layer_sizes(input_size, btl_size, n_layers - 1) + [btl_size]
    dec = [btl_size] + hidden_layer_sizes(btl_size, input_size, n_layers - 1) + [input_size]
    return enc, dec


def synth_state_dict(input_size: int, btl_size: int = 100, n_layers: int = 5, seed: int = 0,
                     enc_out: int | None = None) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the reference's key names.  Linear weights/biases ~ U(+-1/sqrt(K))
    (nn.Linear's range); BN affine and running stats are perturbed away from the
    identity so eval-mode BN is non-trivial: gamma~U(.8,1.2), beta~.1 N, mean~.1 N,
    var~U(.5,1.5).  ``enc_out`` widens the encoder output (VIB: 2*btl)."""
    g = torch.Generator().manual_seed(seed)
    enc_w, dec_w = ae_widths(input_size, btl_size, n_layers)
    if enc_out is not None:
        enc_w = [input_size] + hidden_layer_sizes(input_size, enc_out, n_layers - 1) + [enc_out]
    sd = OrderedDict()
    for prefix, widths in (("encoder", enc_w), ("decoder", dec_w)):
        n = len(widths) - 1
        for i in range(n):
            K, N = widths[i], widths[i + 1]
            bound = 1.0 / (K ** 0.5)
            sd[f"{prefix}.net.{i}.layer.weight"] = (torch.rand(N, K, generator=g) * 2 - 1) * bound
            sd[f"{prefix}.net.{i}.layer.bias"] = (torch.rand(N, generator=g) * 2 - 1) * bound
            if i < n - 1:
                sd[f"{prefix}.net.{i}.bn.weight"] = 0.8 + 0.4 * torch.rand(N, generator=g)
                sd[f"{prefix}.net.{i}.bn.bias"] = 0.1 * torch.randn(N, generator=g)
                sd[f"{prefix}.net.{i}.bn.running_mean"] = 0.1 * torch.randn(N, generator=g)
                sd[f"{prefix}.net.{i}.bn.running_var"] = 0.5 + torch.rand(N, generator=g)
                sd[f"{prefix}.net.{i}.bn.num_batches_tracked"] = torch.tensor(20, dtype=torch.int64)
    return sd


def synth_windows(n: int, input_size: int, seed: int, anomaly_rate: float = 0.1,
                  label_seed: int = 4321) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x [n,D] fp32 in [0,1], y [n] bool).  Anomalous rows: force-torque slice x0.3 and
    +0.15 N(0,1) on the hand-camera slice, clamped.  For other D
    the slices scale proportionally."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, input_size, generator=g)
    gl = torch.Generator().manual_seed(label_seed + seed)
    y = torch.rand(n, generator=gl) < anomaly_rate
    if y.any():
        cam_hi = (input_size * 1024) // 1728
        ft_lo, ft_hi = (input_size * 1536) // 1728, (input_size * 1600) // 1728
        idx = y.nonzero().squeeze(1)
        noise = 0.15 * torch.randn(idx.numel(), max(cam_hi, 1), generator=g)
        xa = x[idx]
        xa[:, :max(cam_hi, 1)] += noise
        if ft_hi > ft_lo:
            xa[:, ft_lo:ft_hi] *= 0.3
        x[idx] = xa.clamp_(0.0, 1.0)
    return x, y
