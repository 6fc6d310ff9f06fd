"""Small stand-alone device ops of the path (all through libmmad; no torch arithmetic)."""
from __future__ import annotations

import torch

from . import _lib


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.MmadError("CUDA tensors required (no CPU path)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def mse_sum(y_hat: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """modules/loss.py:31-32,47-52: nn.MSELoss(reduction='sum') -> 0-dim device tensor."""
    _need_cuda(y_hat, y)
    a = y_hat.detach().float().contiguous()
    b = y.detach().float().contiguous()
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    out = torch.zeros(1, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().mmad_sq_diff_sum(a.data_ptr(), b.data_ptr(), a.numel(), out.data_ptr(), _stream()))
    return out[0]


def row_mean_sq(d: torch.Tensor) -> torch.Tensor:
    """utils/metric.py:133,171: (d**2).mean(axis=1) for a [n, cols] device matrix."""
    _need_cuda(d)
    d = d.float()
    if d.stride(1) != 1:
        d = d.contiguous()
    n, cols = d.shape
    out = torch.empty(n, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.lib().mmad_row_mean_sq(d.data_ptr(), d.stride(0) if n > 1 else cols, n, cols,
                                               out.data_ptr(), _stream()))
    return out


def vib_reparameterize(output: torch.Tensor, k: int = 1, stochastic: bool = True, eps: torch.Tensor | None = None):
    """decorators/variational_info_bottleneck.py:19-42 ('normal'): split (mu, logvar),
    z[k,B,h] = eps * exp(logvar/2) + mu, or mu broadcast when not stochastic (and grad disabled)."""
    if k < 1:
        raise ValueError("k should be >= 1")
    _need_cuda(output, eps)
    shp = output.shape
    o2 = output.detach().reshape(-1, shp[-1]).float().contiguous()
    B, two_h = o2.shape
    h = two_h // 2
    use_noise = torch.is_grad_enabled() or stochastic
    if use_noise and eps is None:
        eps = torch.randn(k, *shp[:-1], h, device=output.device, dtype=torch.float32)
    z = torch.empty(k, B, h, dtype=torch.float32, device=output.device)
    mu = torch.empty(B, h, dtype=torch.float32, device=output.device)
    logvar = torch.empty(B, h, dtype=torch.float32, device=output.device)
    e = eps.reshape(k, B, h).float().contiguous() if use_noise else None
    with torch.cuda.device(output.device):
        _lib.check(_lib.lib().mmad_vib_reparam(o2.data_ptr(), two_h, B, h, k, e.data_ptr() if e is not None else None,
                                               z.data_ptr(), mu.data_ptr(), logvar.data_ptr(), _stream()))
    lead = shp[:-1]
    return {"z": z.reshape(k, *lead, h), "mu": mu.reshape(*lead, h), "logvar": logvar.reshape(*lead, h)}
