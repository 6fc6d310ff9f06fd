"""Fused training step: ``AutoEncoder.step`` of the reference (models/auto_encoder.py:57-77) with the
forward, loss and backward executed by libmmad (``mmad_train_fwd_bwd``).

``fused_train_loss(model, x)`` returns a 0-dim loss tensor attached to the autograd graph through a
custom Function whose backward hands out the gradients the fused kernels already produced, so the
reference's own step body -- ``loss.backward(); optimizer.step()`` -- works verbatim with any
``torch.optim`` optimizer, and with ``icra2021_multimodal_ad_b200.optim.Adam`` (one multi-tensor launch).

Gradients live in ONE flat fp32 buffer owned by the model (views per parameter, ``parameters()`` order):
data-parallel training all-reduces that buffer once per step (SUM: the loss is a sum, model_builder.py:42).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import TrainLayer, check, lib


def _module_layers(module) -> List:
    return list(module.layer_list)


class TrainState:
    """Per-model device state of the fused step: flat gradient buffer, workspace, pointer tables."""

    def __init__(self, model):
        self.params = list(model.parameters())
        dev = self.params[0].device
        pad4 = lambda k: (k + 3) // 4 * 4  # noqa: E731   every view starts 16-byte aligned (vector loads in Adam)
        n = sum(pad4(p.numel()) for p in self.params)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            off += pad4(p.numel())
        self.view_of = {id(p): v for p, v in zip(self.params, self.views)}
        self.ws: Optional[torch.Tensor] = None
        self.ws_batch = 0
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.anchor = torch.zeros((), dtype=torch.float32, device=dev, requires_grad=True)   # ties the loss to autograd
        self._tables = None
        self._tables_key = None
        self.momentum = next((fc.bn.momentum for fc in _module_layers(model.encoder) if fc.bn is not None), 0.1)
        self.group = None            # torch.distributed group for BatchNorm statistics (SyncBN semantics)
        self.world = 1
        self._cb = None

    def tables(self, model):
        # raw-pointer tables are cached; storage only moves on .to()/.cuda(), which moves every tensor
        key = (self.params[0].data_ptr(), self.params[-1].data_ptr(), self.flat_grad.data_ptr())
        if self._tables is not None and self._tables_key == key:
            return self._tables

        def table(module):
            layers = _module_layers(module)
            arr = (TrainLayer * len(layers))()
            for i, fc in enumerate(layers):
                t = arr[i]
                t.W, t.b = fc.layer.weight.data_ptr(), fc.layer.bias.data_ptr()
                t.gW, t.gb = self.view_of[id(fc.layer.weight)].data_ptr(), self.view_of[id(fc.layer.bias)].data_ptr()
                if fc.bn is not None:
                    t.gamma, t.beta = fc.bn.weight.data_ptr(), fc.bn.bias.data_ptr()
                    t.run_mean, t.run_var = fc.bn.running_mean.data_ptr(), fc.bn.running_var.data_ptr()
                    t.num_batches_tracked = fc.bn.num_batches_tracked.data_ptr()
                    t.ggamma, t.gbeta = self.view_of[id(fc.bn.weight)].data_ptr(), self.view_of[id(fc.bn.bias)].data_ptr()
            return arr
        self._tables, self._tables_key = (table(model.encoder), table(model.decoder)), key
        return self._tables

    def workspace(self, h, batch: int) -> torch.Tensor:
        if self.ws is None or batch > self.ws_batch:
            nbytes = lib().mmad_train_workspace_bytes(h, batch)
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.flat_grad.device)
            self.ws_batch = batch
        return self.ws

    def allreduce_callback(self):
        """C callback handed to mmad_train_fwd_bwd: SUM-all-reduce of a BatchNorm statistics slice of
        the workspace over ``self.group`` (NCCL, enqueued on the current stream)."""
        if self.world <= 1 or getattr(self, "native", False):     # native: the library's own communicator does it
            if self._cb is None or getattr(self, "_cb_world", 1) != 1:
                self._cb_world = 1
                self._cb = None
            if self._cb is None:
                self._cb = _lib.ALLREDUCE_FN(0)
            return self._cb, None
        import torch.distributed as dist
        state = self

        if self._cb is not None and getattr(self, "_cb_world", 1) == self.world:
            return self._cb, None

        def cb(ctx, d_buf, count, stream):
            try:
                off = d_buf - state.ws.data_ptr()
                view = state.ws[off:off + 8 * count].view(torch.float64)
                dist.all_reduce(view, group=state.group)
                return 0
            except Exception:      # never let an exception cross the C boundary
                return -1
        self._cb = _lib.ALLREDUCE_FN(cb)
        self._cb_world = self.world
        return self._cb, None


def train_state(model) -> TrainState:
    st = getattr(model, "_train_state", None)
    if st is not None:       # cheap validity check on the hot path: same first/last Parameter objects, same device
        first, last = st.params[0], st.params[-1]
        enc0 = model.encoder.layer_list[0].layer.weight
        dec_last = model.decoder.layer_list[-1].layer.bias
        if first is enc0 and last is dec_last and st.flat_grad.device == enc0.device:
            return st
    st = TrainState(model)
    model._train_state = st
    return st


def set_data_parallel(model, group=None, native: Optional[bool] = None, overlap_grads: bool = False, peer: Optional[bool] = None):
    """Make BatchNorm batch statistics global over ``group`` (N-GPU data parallel == 1 GPU on the
    concatenated batch); gradients are combined by ``allreduce_gradients``.

    native (default: env MMAD_PY_ALLREDUCE != 1): the library opens its own NCCL communicator
    (``mmad_comm_init``; the unique id travels through ``torch.distributed``) and enqueues the collectives
    itself, so the data-parallel step is one CUDA-graph replay.  Otherwise ``torch.distributed.all_reduce`` is
    called back from inside the step (works with any backend).

    peer (default: env MMAD_NO_PEER != 1): both collectives of the step run over NVLink peer memory (``mmad_peer_*``: every
    rank maps every other rank's buffers through cudaIpc) instead of NCCL -- the 16 BatchNorm-statistics exchanges
    (<= 2 x 1408 doubles each, strictly serialised with the layer chain) INSIDE the one-kernel BatchNorm forward / backward
    (batches <= 512 rows; separate one-kernel exchanges above that), and the flat gradient all-reduce as one kernel per rank
    (the gradient buffer then lives in library-owned, peer-mapped memory; env MMAD_NO_PEER_GRADS=1 keeps it in torch memory
    and on NCCL).  Falls back to NCCL if the GPUs cannot map each other.
    overlap_grads (default off): the gradients are all-reduced inside the captured step in two buckets on the second
    stream, the decoder's while the encoder's backward pass still runs; ``allreduce_gradients`` is then a no-op.  Measured
    at B = 256 per GPU at the start of round 2 (scripts/dp_ablation.py): no better than ONE flat all-reduce after the step at
    2 GPUs (0.987 vs 0.998 ms) and worse at 8 (1.232 vs 1.168 ms) -- the overlapped collective takes SMs from a chain of
    small GEMMs.  (The shipped default now runs the step in 0.55 / 0.63 ms at 2 / 8 GPUs)"""
    import os
    import torch.distributed as dist
    st = train_state(model)
    if getattr(st, "peer_grads", False):
        # a previous call moved the flat gradient buffer into library-owned memory, which the calls below free or replace:
        # back to a torch-owned buffer first, so that nothing ever points at freed memory
        _own_gradient_buffer(st)
    st.group = group
    st.world = dist.get_world_size(group) if dist.is_initialized() else 1
    st.native = False
    if native is None:
        native = os.environ.get("MMAD_PY_ALLREDUCE", "0") != "1"
    if st.world > 1 and native and dist.get_backend(group) == "nccl":
        eng = model.handle_engine()
        dev = eng.device
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if dist.get_rank(group) == 0:
            buf = (C.c_ubyte * 128)()
            check(lib().mmad_comm_unique_id(buf))
            uid.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(uid, src=src, group=group)
        raw = (C.c_ubyte * 128)(*uid.cpu().tolist())
        with torch.cuda.device(dev):
            check(lib().mmad_comm_init(eng._h, raw, dist.get_rank(group), st.world))
            check(lib().mmad_comm_set_grad_allreduce(eng._h, 1 if overlap_grads else 0))
            if peer is None:
                peer = os.environ.get("MMAD_NO_PEER", "0") != "1"
            st.peer = False
            if peer:
                hb = (C.c_ubyte * 64)()
                ok = torch.ones(1, dtype=torch.int32, device=dev)
                try:
                    check(lib().mmad_peer_create(eng._h, hb))
                except _lib.MmadError:
                    ok.zero_()
                mine = torch.tensor(list(hb), dtype=torch.uint8, device=dev)
                allh = [torch.empty_like(mine) for _ in range(st.world)]
                dist.all_gather(allh, mine, group=group)
                if int(ok.item()):
                    try:
                        raw = (C.c_ubyte * (64 * st.world))(*torch.cat(allh).cpu().tolist())
                        check(lib().mmad_peer_open(eng._h, raw, dist.get_rank(group), st.world))
                    except _lib.MmadError:
                        ok.zero_()
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # all ranks or none (a mixed set would deadlock)
                if int(ok.item()):
                    st.peer = True
                else:
                    check(lib().mmad_peer_close(eng._h))
                dist.barrier(group)
                if st.peer and os.environ.get("MMAD_NO_PEER_GRADS", "0") != "1":
                    _peer_gradient_buffer(st, eng, group, dev)
        st.native = True
        st.native_handle = eng._h.value      # the communicator lives in THIS library handle
        st.grads_in_step = bool(overlap_grads)
    st.batch_checked = None
    return st


def _rebind_gradient_views(st, flat):
    st.flat_grad = flat
    st.views, off = [], 0
    pad4 = lambda k: (k + 3) // 4 * 4  # noqa: E731
    for p in st.params:
        st.views.append(flat[off:off + p.numel()].view_as(p))
        off += pad4(p.numel())
        p.grad = None
    st.view_of = {id(p): v for p, v in zip(st.params, st.views)}
    st._tables = None


def _own_gradient_buffer(st):
    _rebind_gradient_views(st, torch.zeros(st.flat_grad.numel(), dtype=torch.float32, device=st.params[0].device))
    st.peer_grads = False


class _DevicePtr:
    """A library-owned device buffer as seen by ``torch.as_tensor`` (zero copy; torch keeps this object alive)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _peer_gradient_buffer(st, eng, group, dev):
    """Move the flat gradient buffer into memory the library allocated and every peer mapped (``mmad_peer_grad_alloc`` /
    ``mmad_peer_grad_open``): ``allreduce_gradients`` then runs as ONE kernel per rank over NVLink peer memory instead of
    ncclAllReduce (include/mmad.h).  All ranks or none."""
    import torch.distributed as dist
    n = st.flat_grad.numel()
    hb = (C.c_ubyte * 64)()
    ptr = C.c_void_p()
    ok = torch.ones(1, dtype=torch.int32, device=dev)
    try:
        check(lib().mmad_peer_grad_alloc(eng._h, n, C.byref(ptr), hb))
    except _lib.MmadError:
        ok.zero_()
    mine = torch.tensor(list(hb), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(st.world)]
    dist.all_gather(allh, mine, group=group)
    if int(ok.item()):
        try:
            raw = (C.c_ubyte * (64 * st.world))(*torch.cat(allh).cpu().tolist())
            check(lib().mmad_peer_grad_open(eng._h, raw))
        except _lib.MmadError:
            ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    dist.barrier(group)
    st.peer_grads = bool(int(ok.item()))
    if not st.peer_grads:
        return
    n4 = (n + 3) // 4 * 4
    flat = torch.as_tensor(_DevicePtr(ptr.value, n4), device=dev)
    assert flat.data_ptr() == ptr.value
    _rebind_gradient_views(st, flat)


def _check_native_comm(st, eng):
    """The library-owned communicator dies with its handle: if the Engine was recreated (device move) after
    set_data_parallel, the collectives would silently become no-ops while the statistics are still divided by the
    global batch."""
    if getattr(st, "native", False) and st.world > 1 and getattr(st, "native_handle", None) != eng._h.value:
        raise _lib.MmadError("the model's device engine was recreated after set_data_parallel(): call "
                             "set_data_parallel(model, group) again")


def _check_equal_batches(st, B: int):
    """BatchNorm statistics and the unbiased running variance are divided by B * world: every rank must hold the same
    number of rows.  Checked once per batch size (one tiny all-reduce), not per step."""
    if st.world <= 1 or getattr(st, "batch_checked", None) == B:
        return
    import torch.distributed as dist
    t = torch.tensor([B, -B], dtype=torch.int64, device=st.flat_grad.device if dist.get_backend(st.group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=st.group)
    if int(t[0]) != B or int(-t[1]) != B:
        raise ValueError(f"data-parallel step: local batches differ across ranks (this rank {B}, min {int(-t[1])}, "
                         f"max {int(t[0])}); pad or drop the last shard so every rank holds the same number of rows")
    st.batch_checked = B


def allreduce_gradients(model):
    """SUM-all-reduce of the flat gradient buffer (one collective per step)."""
    import torch.distributed as dist
    st = train_state(model)
    if st.world > 1:
        if getattr(st, "native", False):
            if getattr(st, "grads_in_step", False):
                return      # already all-reduced per layer inside the captured step (mmad_comm_set_grad_allreduce)
            eng = model.handle_engine()
            _check_native_comm(st, eng)
            with torch.cuda.device(eng.device):
                check(lib().mmad_comm_allreduce_f32(eng._h, st.flat_grad.data_ptr(), st.flat_grad.numel(),
                                                    torch.cuda.current_stream().cuda_stream))
        else:
            dist.all_reduce(st.flat_grad, group=st.group)


class _FusedStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, eps, beta_kl, anchor):
        st = train_state(model)
        eng = model.handle_engine()
        B = x.shape[0]
        if st.world > 1:
            _check_native_comm(st, eng)
            _check_equal_batches(st, B)
        ws = st.workspace(eng._h, B)
        enc_t, dec_t = st.tables(model)
        cb, _ = st.allreduce_callback()
        momentum = st.momentum
        with torch.cuda.device(x.device):
            check(lib().mmad_train_fwd_bwd(eng._h, x.data_ptr(), x.stride(0) if B > 1 else x.shape[1], B, B * st.world,
                                           enc_t, dec_t, eps.data_ptr() if eps is not None else None, float(beta_kl),
                                           float(momentum), st.loss.data_ptr(), ws.data_ptr(), ws.numel(), cb, None,
                                           torch.cuda.current_stream().cuda_stream))
        model._train_steps = getattr(model, "_train_steps", 0) + 1    # parameters/buffers changed under raw pointers
        ctx.model = model
        return st.loss[0].clone()

    @staticmethod
    def backward(ctx, gout):
        st = train_state(ctx.model)
        flat = st.flat_grad
        # d(loss)/d(loss) scaling.  loss.backward() -- the reference's step body -- seeds autograd with ones; only a
        # caller-supplied seed (loss.backward(g), a scaled loss) needs the multiply.  The kernel reads the seed on the
        # device and returns at once when it is 1 (this used to rewrite 41 MB per step to multiply by 1.0)
        g = gout if (gout.dtype == torch.float32 and gout.is_contiguous()) else gout.float().contiguous()
        check(lib().mmad_scale_unless_one(flat.data_ptr(), flat.numel(), g.data_ptr(), torch.cuda.current_stream().cuda_stream))
        # hand the gradients over directly (36 AccumulateGrad nodes cost ~0.7 ms of host time per step)
        for p, v in zip(st.params, st.views):
            if p.grad is None:
                p.grad = v
            else:
                p.grad.add_(v)
        return None, None, None, None, None


def fused_train_loss(model, x: torch.Tensor, eps: Optional[torch.Tensor] = None, beta_kl: float = 0.0) -> torch.Tensor:
    if not x.is_cuda:
        raise _lib.MmadError("training needs CUDA tensors (no CPU path)")
    x = x.detach().float()
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    if eps is not None:
        eps = eps.detach().to(x.device, torch.float32).reshape(x.shape[0], -1).contiguous()
    st = train_state(model)
    if any(p.grad is not None for p in st.params):
        for p, v in zip(st.params, st.views):
            # a .grad left over from the previous step aliases the flat buffer this step overwrites: give it its
            # own storage so autograd's accumulation (zero_grad(set_to_none=False) callers) stays correct
            if p.grad is not None and p.grad.data_ptr() == v.data_ptr():
                p.grad = p.grad.clone()
    out = _FusedStep.apply(model, x, eps, beta_kl, st.anchor)
    st.last_loss = out          # step_loss() recognises the tensor the step produced
    return out


def step_loss(model, loss: torch.Tensor) -> float:
    """``float(loss)`` for the loss tensor ``fused_train_loss`` just returned, WITHOUT waiting for the backward pass, the
    optimizer and whatever else is queued on the stream: the step publishes the value to mapped pinned memory as soon as
    its forward pass has it (``mmad_train_loss``).  ``AutoEncoder.step`` returns the loss every step
    (models/auto_encoder.py:77); read through the stream, the host would only get back to launching the next step after the
    GPU has gone idle.  Any other tensor (a scaled loss, a validation loss) takes the ordinary ``float()``."""
    st = getattr(model, "_train_state", None)
    if st is None or getattr(st, "last_loss", None) is not loss:
        return float(loss.detach())
    out = C.c_float()
    if lib().mmad_train_loss(model.handle_engine()._h, C.byref(out)) != 0:
        return float(loss.detach())      # nothing was published (the step ran inside a caller's own stream capture)
    return float(out.value)
