"""Device engine: one ``mmad_t`` handle (packed weights + NAP fit) bound to a CUDA device.

Thin host-side plumbing over the C ABI: torch owns device memory and streams, libmmad
does the arithmetic.  Nothing on the scoring path computes on the CPU or through torch ops; the NAP FIT (off the
timed path) calls two library factorizations, ``torch.linalg.eigh`` and ``torch.linalg.qr`` (cuSOLVER), on the
D' x D' Gram matrix.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Desc, PREC, check, lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def clamp_layer_range(n_diffs: int, start: int, end: Optional[int]):
    """utils/metric.py:155-162 followed by python slice semantics ``diffs[start:end]``."""
    if end is None:
        end = n_diffs + 1
    if start > n_diffs - 1:
        start = n_diffs - 1
    if end - start < 1:
        end = start + 1
    lo, hi, _ = slice(start, end).indices(n_diffs)
    return lo, hi


class Engine:
    def __init__(self, enc_widths: Sequence[int], dec_widths: Sequence[int], precision: str = "fp32",
                 device: Optional[torch.device] = None, lrelu_slope: float = 0.2, bn_eps: float = 1e-5):
        if not torch.cuda.is_available():
            raise _lib.MmadError("icra2021_multimodal_ad_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.enc_widths, self.dec_widths = list(enc_widths), list(dec_widths)
        d = Desc()
        d.n_enc, d.n_dec = len(enc_widths) - 1, len(dec_widths) - 1
        for i, w in enumerate(enc_widths):
            d.enc_widths[i] = w
        for i, w in enumerate(dec_widths):
            d.dec_widths[i] = w
        d.lrelu_slope, d.bn_eps, d.precision = lrelu_slope, bn_eps, PREC[precision]
        self.precision = precision
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().mmad_create(C.byref(d), C.byref(self._h)))
        self._ws: Optional[torch.Tensor] = None
        self._ws_rows = 0
        self.nap_range = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib().mmad_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------------
    @property
    def D(self) -> int:
        return self.enc_widths[0]

    @property
    def n_diffs(self) -> int:
        return len(self.enc_widths)

    def set_precision(self, precision: str):
        check(lib().mmad_set_precision(self._h, PREC[precision]))
        if precision != self.precision:
            self.nap_range = None        # the library dropped the fit: its variances belong to the old arithmetic
            self.nap_fit_state = None
        self.precision = precision
        self._ws = None
        self._ws_rows = 0

    def set_option(self, name: str, value: float):
        """``mmad_set_option``: 'acc_comp', 'nap_passes', 'require_pinned', 'smallnet' (include/mmad.h).  Changing
        'nap_passes' drops the installed NAP fit (its variances belong to the old arithmetic): refit afterwards."""
        check(lib().mmad_set_option(self._h, name.encode(), float(value)))
        if name == "nap_passes":
            self.nap_range = None
            self.nap_fit_state = None

    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Pack a reference-format state dict (keys ``encoder.net.{i}.layer.weight`` ...).  An installed NAP fit
        belongs to the old weights and is dropped (the reference refits on every test() call)."""
        self.nap_range = None
        self.nap_fit_state = None
        with torch.cuda.device(self.device):
            for m, prefix, widths in ((0, "encoder", self.enc_widths), (1, "decoder", self.dec_widths)):
                for i in range(len(widths) - 1):
                    def g(name):
                        t = sd.get(f"{prefix}.net.{i}.{name}")
                        if t is None:
                            return None
                        return t.detach().to(self.device, torch.float32).contiguous()
                    W, b = g("layer.weight"), g("layer.bias")
                    if W is None or tuple(W.shape) != (widths[i + 1], widths[i]):
                        raise ValueError(f"{prefix}.net.{i}.layer.weight missing or wrong shape")
                    bn = [g("bn.weight"), g("bn.bias"), g("bn.running_mean"), g("bn.running_var")]
                    check(lib().mmad_set_layer(self._h, m, i, _ptr(W), _ptr(b), *[_ptr(t) for t in bn], _stream()))
            torch.cuda.current_stream().synchronize()   # the temporaries above may be freed now

    # ---- workspace ---------------------------------------------------------------------
    def workspace(self, rows: int) -> torch.Tensor:
        rows = max(128, min(int(rows), 4 * 148 * 128))
        if self._ws is None or rows > self._ws_rows:
            nbytes = lib().mmad_workspace_bytes(self._h, rows)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_rows = rows
        return self._ws

    def _check_x(self, x: torch.Tensor) -> torch.Tensor:
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise _lib.MmadError("input must be a CUDA tensor (no CPU path)")
        if x.dim() != 2 or x.shape[1] != self.D:
            raise ValueError(f"expected [n, {self.D}] input, got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            x = x.float()
        if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < self.D):
            x = x.contiguous()
        return x

    # ---- forward / loss ------------------------------------------------------------------
    def forward(self, x: torch.Tensor, want_code: bool = False):
        x = self._check_x(x)
        n = x.shape[0]
        xhat = torch.empty(n, self.D, dtype=torch.float32, device=x.device)
        z = torch.empty(n, self.enc_widths[-1], dtype=torch.float32, device=x.device) if want_code else None
        ws = self.workspace(n)
        with torch.cuda.device(self.device):
            check(lib().mmad_ae_forward(self._h, x.data_ptr(), x.stride(0) if n > 1 else self.D, n, xhat.data_ptr(),
                                        _ptr(z), ws.data_ptr(), ws.numel(), _stream()))
        return (xhat, z) if want_code else xhat

    def recon_loss(self, x: torch.Tensor) -> torch.Tensor:
        x = self._check_x(x)
        n = x.shape[0]
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        ws = self.workspace(n)
        with torch.cuda.device(self.device):
            check(lib().mmad_recon_loss(self._h, x.data_ptr(), x.stride(0) if n > 1 else self.D, n, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), _stream()))
        return out[0]

    # ---- scoring ------------------------------------------------------------------------
    def concat_width(self, lo: int, hi: int) -> int:
        r = lib().mmad_concat_width(self._h, lo, hi)
        if r < 0:
            check(r)
        return r

    def score(self, x: torch.Tensor, lo: int = 0, hi: Optional[int] = None, base: bool = True, sap: bool = True,
              nap: bool = False, diffs: bool = False) -> Dict[str, torch.Tensor]:
        """Fused get_diffs + base/SAP/NAP scores for rows of ``x`` (device tensor)."""
        x = self._check_x(x)
        n = x.shape[0]
        hi = self.n_diffs if hi is None else hi
        out: Dict[str, torch.Tensor] = {}
        mk = lambda: torch.empty(n, dtype=torch.float32, device=x.device)  # noqa: E731
        if base:
            out["base"] = mk()
        if sap:
            out["sap"] = mk()
        if nap:
            out["nap"] = mk()
        if diffs:
            out["diffs"] = torch.empty(n, self.concat_width(lo, hi), dtype=torch.float32, device=x.device)
        ws = self.workspace(n)
        with torch.cuda.device(self.device):
            check(lib().mmad_score(self._h, x.data_ptr(), x.stride(0) if n > 1 else self.D, n, lo, hi,
                                   _ptr(out.get("base")), _ptr(out.get("sap")), _ptr(out.get("nap")),
                                   _ptr(out.get("diffs")), ws.data_ptr(), ws.numel(), _stream()))
        return out

    def score_host(self, x: np.ndarray, lo: int = 0, hi: Optional[int] = None, base: bool = True, sap: bool = True,
                   nap: bool = False) -> Dict[str, np.ndarray]:
        """Host-buffer entry point (``mmad_score_host``): x is a C-contiguous fp32 ndarray (ideally in
        pinned memory); returns host arrays.  H2D/compute/D2H are pipelined inside the library; calls of <= 16 rows
        without NAP are one kernel launch (the realtime path: keep this wrapper thin, it is ~12 us of a 62-us call)."""
        if isinstance(x, torch.Tensor):
            x = x.numpy()
        if x.dtype != np.float32 or not x.flags["C_CONTIGUOUS"]:
            x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        if hi is None:
            hi = len(self.enc_widths)
        out = {}
        pb = ps = pn = None
        if base:
            out["base"] = a = np.empty(n, dtype=np.float32); pb = a.ctypes.data
        if sap:
            out["sap"] = a = np.empty(n, dtype=np.float32); ps = a.ctypes.data
        if nap:
            out["nap"] = a = np.empty(n, dtype=np.float32); pn = a.ctypes.data
        if torch.cuda.current_device() == self.device.index:      # realtime calls: skip the device guard (~4 us)
            rc = lib().mmad_score_host(self._h, x.ctypes.data, x.shape[1], n, lo, hi, pb, ps, pn)
        else:
            with torch.cuda.device(self.device):
                rc = lib().mmad_score_host(self._h, x.ctypes.data, x.shape[1], n, lo, hi, pb, ps, pn)
        if rc:
            check(rc)
        return out

    def stream_input(self, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        """``mmad_stream_input``: the pinned, device-mapped ``[64, D]`` input buffer of the one-launch realtime kernel as a
        NumPy array.  ``score_host(buf[:n], ...)`` on a leading slice of it skips the staging copy."""
        hi = self.n_diffs if hi is None else hi
        p, mr = C.POINTER(C.c_float)(), C.c_int()
        with torch.cuda.device(self.device):
            check(lib().mmad_stream_input(self._h, lo, hi, C.byref(p), C.byref(mr)))
        return np.ctypeslib.as_array(p, shape=(mr.value, self.D))

    # ---- NAP fit ------------------------------------------------------------------------
    def nap_accumulate_sum(self, x: torch.Tensor, lo: int, hi: int, acc: torch.Tensor):
        x = self._check_x(x)
        ws = self.workspace(x.shape[0])
        with torch.cuda.device(self.device):
            check(lib().mmad_nap_accumulate_sum(self._h, x.data_ptr(), x.stride(0) if x.shape[0] > 1 else self.D,
                                                x.shape[0], lo, hi, acc.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))

    def nap_accumulate_gram(self, x: torch.Tensor, lo: int, hi: int, mu: torch.Tensor, gram: torch.Tensor):
        x = self._check_x(x)
        ws = self.workspace(x.shape[0])
        with torch.cuda.device(self.device):
            check(lib().mmad_nap_accumulate_gram(self._h, x.data_ptr(), x.stride(0) if x.shape[0] > 1 else self.D,
                                                 x.shape[0], lo, hi, mu.data_ptr(), gram.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), _stream()))

    def nap_rotate_stats(self, x: torch.Tensor, lo: int, hi: int, rsum: torch.Tensor, rsq: torch.Tensor):
        x = self._check_x(x)
        ws = self.workspace(x.shape[0])
        with torch.cuda.device(self.device):
            check(lib().mmad_nap_rotate_stats(self._h, x.data_ptr(), x.stride(0) if x.shape[0] > 1 else self.D,
                                              x.shape[0], lo, hi, rsum.data_ptr(), rsq.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _stream()))

    def nap_set_standardizer(self, var: torch.Tensor, mu2: torch.Tensor):
        var = var.detach().to(self.device, torch.float32).contiguous()
        mu2 = mu2.detach().to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            check(lib().mmad_nap_set_standardizer(self._h, var.data_ptr(), mu2.data_ptr(), _stream()))
            torch.cuda.current_stream().synchronize()

    def nap_set_fit(self, lo: int, hi: int, mu: torch.Tensor, vt: torch.Tensor, var: torch.Tensor, mu2: torch.Tensor):
        f = lambda t: t.detach().to(self.device, torch.float32).contiguous()  # noqa: E731
        mu, vt, var, mu2 = f(mu), f(vt), f(var), f(mu2)
        with torch.cuda.device(self.device):
            check(lib().mmad_nap_set_fit(self._h, lo, hi, vt.shape[0], mu.data_ptr(), vt.data_ptr(), var.data_ptr(),
                                         mu2.data_ptr(), _stream()))
            torch.cuda.current_stream().synchronize()
        self.nap_range = (lo, hi)
        self._ws = None      # NAP slot count may have changed
        self._ws_rows = 0

    def _allreduce(self, t: torch.Tensor, group=None):
        """SUM all-reduce of a device tensor across the ranks: through the library's own communicator when one is
        installed (``mmad_comm_init``, fp32 / fp64), else ``torch.distributed``."""
        import torch.distributed as dist
        if group is None and lib().mmad_comm_world(self._h) > 1 and t.dtype in (torch.float32, torch.float64) and t.is_contiguous():
            fn = lib().mmad_comm_allreduce_f64 if t.dtype == torch.float64 else lib().mmad_comm_allreduce_f32
            with torch.cuda.device(self.device):
                check(fn(self._h, t.data_ptr(), t.numel(), _stream()))
        else:
            dist.all_reduce(t, group=group)

    def nap_fit(self, x_train: torch.Tensor, lo: int = 0, hi: Optional[int] = None, group=None,
                batch_rows: int = 16384, distributed: Optional[bool] = None,
                restandardize: bool = True, factor: str = "hybrid", phases: Optional[dict] = None, tau: float = 1e-5) -> Dict[str, torch.Tensor]:
        """utils/normalize.py:47-70 + 20-34 on device, from statistics instead of an SVD of the
        N x D' matrix:  mu = mean(d);  G = (d-mu)^T (d-mu)  (fp64)  = V diag(lambda) V^T;
        var_j = lambda_j / (N-1)  (== diag(np.cov) of the rotated data); K = min(N, D').
        ``distributed`` (default: whether torch.distributed is initialised) / ``group``: the row
        shards' sum and Gram are all-reduced (one exchange per pass; the Gram exchange moves the
        upper triangle only).  mu2 (the Standardizer mean of the rotated train data) is identically zero in exact
        arithmetic and is installed as zero.  ``phases`` (a dict) receives wall-clock seconds per phase (adds syncs)."""
        import time
        import torch.distributed as dist
        use_dist = (dist.is_available() and dist.is_initialized()) if distributed is None else bool(distributed)
        hi = self.n_diffs if hi is None else hi
        dsel = self.concat_width(lo, hi)
        dev = self.device
        n_local = x_train.shape[0]
        t_last = [time.perf_counter()]

        def mark(name):
            if phases is not None:
                torch.cuda.synchronize(dev)
                now = time.perf_counter()
                phases[name] = phases.get(name, 0.0) + now - t_last[0]
                t_last[0] = now
        if phases is not None:
            torch.cuda.synchronize(dev)
            t_last[0] = time.perf_counter()
        s = torch.zeros(dsel + 1, dtype=torch.float64, device=dev)      # column sums and, last, the row count
        for r0 in range(0, n_local, batch_rows):
            self.nap_accumulate_sum(x_train[r0:r0 + batch_rows], lo, hi, s)
        s[dsel] = n_local
        mark("chain_sum_s")
        if use_dist:
            self._allreduce(s, group)
        N = int(round(s[dsel].item()))
        mu = (s[:dsel] / N).float()
        mark("exchange_s")
        gram = torch.zeros(dsel, dsel, dtype=torch.float64, device=dev)
        for r0 in range(0, n_local, batch_rows):
            self.nap_accumulate_gram(x_train[r0:r0 + batch_rows], lo, hi, mu, gram)
        mark("chain_gram_s")
        if use_dist:
            tri = torch.empty(dsel * (dsel + 1) // 2, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                check(lib().mmad_tri_pack(gram.data_ptr(), dsel, tri.data_ptr(), _stream()))
            self._allreduce(tri, group)
            with torch.cuda.device(dev):
                check(lib().mmad_tri_unpack(tri.data_ptr(), dsel, gram.data_ptr(), _stream()))
            del tri
        mark("exchange_s")
        fit = nap_fit_from_stats(mu, gram, N, factor=factor, tau=tau)
        del gram
        mark("eig_factor_s")
        self.nap_set_fit(lo, hi, fit["mu"], fit["vt"], fit["var"], fit["mu2"])
        check(lib().mmad_nap_set_structure(self._h, int(fit["tri_rows"])))
        mark("pack_s")
        if restandardize:
            # Standardizer.fit on Rotater.run(train) (utils/metric.py:214-216): third pass, the rotation done
            # by the scoring kernels themselves so their rounding noise in near-null directions
            # is normalised exactly like the reference normalises its own
            K = fit["vt"].shape[0]
            rs = torch.zeros(2, K, dtype=torch.float64, device=dev)
            for r0 in range(0, n_local, batch_rows):
                self.nap_rotate_stats(x_train[r0:r0 + batch_rows], lo, hi, rs[0], rs[1])
            mark("chain_rotate_s")
            if use_dist:
                self._allreduce(rs, group)
            mu2 = rs[0] / N
            var = (rs[1] - N * mu2 * mu2) / (N - 1)
            fit["mu2"], fit["var_eig"], fit["var"] = mu2.float(), fit["var"], var.float()
            self.nap_set_standardizer(fit["var"], fit["mu2"])
            mark("exchange_s")
        fit["lo"], fit["hi"], fit["precision"] = lo, hi, self.precision
        self.nap_fit_state = fit
        return fit

    # ---- NAP-fit checkpoint: (mu, factor rows, var, mu2, N) instead of the raw N x D' train diffs ----
    def nap_state_dict(self) -> Dict[str, object]:
        """The installed NAP fit as a compact artefact -- what replaces ``torch.save(train_diffs, config.train_diffs)``
        (utils/metric.py:205; test_file/FullTest.py:33-44 re-runs the SVD from those diffs on every call)."""
        f = getattr(self, "nap_fit_state", None)
        if f is None or self.nap_range is None:
            raise _lib.MmadError("no NAP fit installed")
        return {"format": "mmad-nap-fit-1", "lo": f["lo"], "hi": f["hi"], "n": f["n"], "factor": f["factor"], "tri_rows": int(f["tri_rows"]),
                "precision": f["precision"], "enc_widths": list(self.enc_widths),
                "mu": f["mu"].cpu(), "vt": f["vt"].cpu(), "var": f["var"].cpu(), "mu2": f["mu2"].cpu()}

    def load_nap_state_dict(self, st: Dict[str, object]):
        if st.get("format") != "mmad-nap-fit-1":
            raise ValueError("not a NAP-fit checkpoint")
        if list(st["enc_widths"]) != list(self.enc_widths):
            raise ValueError("NAP-fit checkpoint belongs to a model with other widths")
        if st["precision"] != self.precision:
            raise ValueError(f"NAP fit was made in {st['precision']} arithmetic, the engine runs {self.precision}: refit "
                             "(the variances of near-null directions carry the mode's rounding noise)")
        self.nap_set_fit(st["lo"], st["hi"], st["mu"], st["vt"], st["var"], st["mu2"])
        check(lib().mmad_nap_set_structure(self._h, int(st["tri_rows"])))
        self.nap_fit_state = {k: st[k] for k in ("mu", "vt", "var", "mu2", "n", "factor", "tri_rows", "lo", "hi", "precision")}


def nap_fit_from_stats(mu: torch.Tensor, gram: torch.Tensor, n_total: int, factor: str = "hybrid", tau: float = 1e-5) -> Dict[str, torch.Tensor]:
    """Eigendecomposition of the centred Gram matrix (fp64, cuSOLVER syevd through
    torch.linalg.eigh -- a library call, not on the hot path) -> (mu, V^T, var, mu2).

    factor="eigen": rows of ``vt`` are the right singular vectors v_j like the reference's Rotater
    (utils/normalize.py:67) and ``var`` their variances.
    factor="triangular": the same score sum_j ((d-mu).v_j)^2 / var_j = |R (d-mu)|^2 through the upper
    triangular factor R of the whitening matrix diag(var^-1/2) V^T = Q R (fp64 Householder QR): half of R is
    zero, so the scoring GEMM does half the products.  Rows are normalised to unit max (the scale goes into
    ``var``) so they split cleanly into fp16 pairs.
    factor="hybrid" (default): triangular factor of the part of the spectrum ABOVE the rounding-noise floor (singular
    values >= tau * the largest, tau = 1e-5) followed by the plain eigenvector rows of the directions below it.  Every
    row of a triangular factor carries the gain of the weakest direction it spans, so on a rank-deficient selection
    (all layers, SURVEY F5: 770 of the 5482 singular values of a random-init model, 1570 of a trained one, sit at the
    fp32 noise floor, 1e-8 of the largest) the noise-floor directions' gain lands in EVERY output; the eigenvector form
    confines it to their own outputs, whose variance the Standardizer refit then normalises like the reference does.
    Measured at D = 1728, all layers, f16x3 (scripts/nap_tau_sweep.py; rank agreement with the fp64 value / AUROC):
    reference golden model -- reference 0.974 / 0.760, eigen = hybrid at every tau in [1e-5, 1e-2] 0.970 / 0.755,
    triangular 0.861 / 0.777; trained model -- reference algorithm 0.954 / 0.794, eigen = hybrid 0.943 / 0.795,
    triangular 0.925 / 0.775.  Well-conditioned selections have nothing below the floor and keep the full triangular
    factor (half the MMA work)."""
    lam, V = torch.linalg.eigh(gram)            # ascending
    # near-null directions can come out slightly negative; floor at fp64 resolution of the
    # largest eigenvalue so that var stays positive like the reference's np.cov diagonal
    lam = lam.flip(0).clamp_min(lam.max() * 1e-16)
    V = V.flip(1)
    K = min(n_total, gram.shape[0])
    var64 = lam[:K] / (n_total - 1)
    tri_rows = 0
    if factor in ("triangular", "hybrid"):
        ks = K if factor == "triangular" else int((lam[:K] >= (tau * tau) * lam[0]).sum().item())
        W = V[:, :ks].t() / var64[:ks].sqrt().unsqueeze(1)                 # ks x D' whitening matrix of the strong part
        R = torch.linalg.qr(W, mode="r").R                                 # ks x D', upper triangular / trapezoidal
        scale = R.abs().amax(dim=1).clamp_min(1e-300)
        vt = torch.triu(R / scale.unsqueeze(1))
        var = 1.0 / (scale * scale)
        if ks < K:                                                         # weak directions: plain eigenvector rows
            vt = torch.cat([vt, V[:, ks:K].t()], dim=0)
            var = torch.cat([var, var64[ks:K]])
        vt, var, tri_rows = vt.contiguous().float(), var.float(), ks
    elif factor == "eigen":
        vt = V[:, :K].t().contiguous().float()
        var = var64.float()
    else:
        raise ValueError("factor must be 'eigen', 'triangular' or 'hybrid'")
    return {"mu": mu.float(), "vt": vt, "var": var, "mu2": torch.zeros(K, dtype=torch.float32, device=mu.device),
            "n": n_total, "factor": factor, "tri_rows": tri_rows}
