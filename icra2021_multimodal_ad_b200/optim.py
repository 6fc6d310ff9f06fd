"""torch.optim.Adam (the optimizer of novelty_detection.py:90) as ONE multi-tensor launch of libmmad
(``mmad_adam_step``): same update rule and defaults (lr 1e-3, betas (0.9, 0.999), eps 1e-8, no weight
decay, no amsgrad).  Drop-in for ``optim.Adam(model.parameters())``; state lives in flat device buffers."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("weight_decay / amsgrad are not used on the reference path")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}

    def _group_state(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        key = tuple(id(p) for p in ps)
        st = self._tables.get(gi)
        if st is None or st["key"] != key:
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.MmadError("mmad Adam needs contiguous fp32 CUDA parameters (no CPU path)")
            pad4 = lambda k: (k + 3) // 4 * 4  # noqa: E731   16-byte aligned slots (vector loads in the kernel)
            n = sum(pad4(p.numel()) for p in ps)
            dev = ps[0].device
            prev = st
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            v = torch.zeros(n, dtype=torch.float32, device=dev)
            st = {"key": key, "m": m, "v": v, "step": prev["step"] if prev else 0, "params": ps}
            off, mp, vp = 0, [], []
            for p in ps:
                mp.append(m[off:off + p.numel()])
                vp.append(v[off:off + p.numel()])
                self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"] = mp[-1].view_as(p), vp[-1].view_as(p)
                off += pad4(p.numel())
            arr = lambda ptrs: (C.c_void_p * len(ptrs))(*ptrs)  # noqa: E731
            st["P"] = arr([p.data_ptr() for p in ps])
            st["M"] = arr([t.data_ptr() for t in mp])
            st["V"] = arr([t.data_ptr() for t in vp])
            st["numel"] = (C.c_longlong * len(ps))(*[p.numel() for p in ps])
            self._tables[gi] = st
        return st

    def zero_grad(self, set_to_none: bool = True):
        """torch.optim.Optimizer.zero_grad without the profiler scope and foreach bookkeeping (30 us of host time in front
        of every step's launch for 36 parameters)."""
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            st = self._group_state(gi, group)
            ps = st["params"]
            # gradient pointer table: reused while the gradients keep living where they did (views of one flat buffer)
            ga, gb = ps[0].grad, ps[-1].grad
            flat = ga.untyped_storage().data_ptr() == gb.untyped_storage().data_ptr()
            gkey = (ga.data_ptr(), gb.data_ptr()) if flat else None
            if gkey is None or st.get("Gkey") != gkey:
                st["G"] = (C.c_void_p * len(ps))(*[p.grad.data_ptr() for p in ps])
                st["Gkey"] = gkey
            G = st["G"]
            st["step"] += 1
            b1, b2 = group["betas"]
            with torch.cuda.device(ps[0].device):
                _lib.check(_lib.lib().mmad_adam_step(len(ps), st["P"], G, st["M"], st["V"], st["numel"], st["step"],
                                                     float(group["lr"]), float(b1), float(b2), float(group["eps"]), 1.0,
                                                     torch.cuda.current_stream().cuda_stream))
            for p in ps:
                self.state[p]["step"] = st["step"]
        return loss
