"""utils/normalize.py of the reference on device: Standardizer, Rotater, Truncater.

``fit``/``run`` keep the reference signatures (``gpu_id``, ``max_size``); inputs may be numpy arrays
(results come back as numpy, like the reference) or CUDA tensors (``as_tensor=True`` keeps results on
device).  Rotater.fit obtains V from the eigendecomposition of the centred Gram matrix accumulated in
fp64 by libmmad instead of an SVD of the N x D matrix (same right singular vectors; their sign and
order do not affect any score).
"""
import numpy as np
import torch

from .. import _lib

_ws = {}


def _dev():
    if not torch.cuda.is_available():
        raise _lib.MmadError("normalisers need a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _as_dev(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    x = x.detach()
    if not x.is_cuda:
        x = x.to(_dev())
    x = x.float()
    if x.dim() != 2:
        raise ValueError("expected a 2-D matrix")
    if x.stride(1) != 1:
        x = x.contiguous()
    return x


def _workspace(cols, device):
    need = _lib.lib().mmad_normalizer_workspace_bytes(int(cols))
    t = _ws.get(str(device))
    if t is None or t.numel() < need:
        t = torch.empty(need, dtype=torch.uint8, device=device)
        _ws[str(device)] = t
    return t


def _stream():
    return torch.cuda.current_stream().cuda_stream


def col_stats(x, want_var):
    n, cols = x.shape
    mean = torch.empty(cols, dtype=torch.float32, device=x.device)
    var = torch.empty(cols, dtype=torch.float32, device=x.device) if want_var else None
    ws = _workspace(cols, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mmad_col_stats(x.data_ptr(), x.stride(0), n, cols, mean.data_ptr(),
                                             var.data_ptr() if want_var else None, ws.data_ptr(), ws.numel(), _stream()))
    return mean, var


class Standardizer():
    def __init__(self, *args, **kwargs):
        self.mu, self.var = None, None

    def fit(self, x):
        # utils/normalize.py:25-34: mean, then diag(np.cov) (fp64, ddof=1) -> fp32
        x = _as_dev(x)
        self.mu, self.var = col_stats(x, True)

    def run(self, x, as_tensor=False):
        # utils/normalize.py:36-45: (x - mu) / var**.5
        x = _as_dev(x)
        out = torch.empty_like(x, memory_format=torch.contiguous_format)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mmad_standardize(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], self.mu.data_ptr(),
                                                   self.var.data_ptr(), out.data_ptr(), out.stride(0), _stream()))
        return out if as_tensor else out.cpu().numpy()


class Rotater():
    def __init__(self, *args, **kwargs):
        self.mu, self.v = None, None

    def fit(self, x, gpu_id=-1):
        # utils/normalize.py:52-70: mu = mean; V = right singular vectors of (x - mu), [D, min(N, D)]
        x = _as_dev(x)
        n, cols = x.shape
        self.mu, _ = col_stats(x, False)
        gram = torch.zeros(cols, cols, dtype=torch.float64, device=x.device)
        ws = _workspace(cols, x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mmad_gram_accumulate(x.data_ptr(), x.stride(0), n, cols, self.mu.data_ptr(),
                                                       gram.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        lam, vec = torch.linalg.eigh(gram)     # small on-device eigendecomposition (cuSOLVER), ascending
        k = min(n, cols)
        self.v = vec.flip(1)[:, :k].float().contiguous()
        self.s = lam.flip(0)[:k].clamp_min(0).sqrt().float()
        self._vt = self.v.t().contiguous()

    def run(self, x, gpu_id=-1, max_size=20000, as_tensor=False):
        # utils/normalize.py:72-103: (x - mu) @ V  (the reference chunks by max_size; so does the library)
        x = _as_dev(x)
        n, cols = x.shape
        k = self._vt.shape[0]
        out = torch.empty(n, k, dtype=torch.float32, device=x.device)
        ws = _workspace(cols, x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mmad_rotate(x.data_ptr(), x.stride(0), n, cols, self.mu.data_ptr(), self._vt.data_ptr(), k,
                                              out.data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(), _stream()))
        return out if as_tensor else out.cpu().numpy()


class Truncater(Rotater):
    """utils/normalize.py:105-146 -- low-rank reconstruction; never called by the reference path."""

    def run(self, x, trunc, gpu_id=-1, max_size=20000):
        raise NotImplementedError("Truncater has no caller in the reference (SURVEY.md section 2); not on the B200 path")
