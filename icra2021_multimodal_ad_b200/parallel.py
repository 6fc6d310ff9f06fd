"""Multi-GPU plumbing (one process per GPU, torch.distributed).

* Scoring shards by sample: contiguous row ranges per rank, weights and the NAP fit replicated, no
  data-path collective; per-sample scores are gathered once for the metrics.
* The NAP fit has one exchange step per pass (column sums, Gram matrix, rotated sums): ``Engine.nap_fit``
  all-reduces them; ``combine_nap_stats`` is the same combination on host tensors (gloo-testable).
* Training is data parallel: BatchNorm statistics are all-reduced inside the step (hook of
  ``mmad_train_fwd_bwd``), gradients once per step as one flat buffer (SUM: the loss is a sum).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of ``rank``: the first ``n % world`` ranks get one extra row."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    q, r = divmod(n, world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-sample results of the ``shard_range`` partition back into sample order
    (4 B/sample for a score vector).  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    rank, ws = world(group)
    if ws == 1:
        return local
    sizes = [shard_range(n_total, r, ws)[1] - shard_range(n_total, r, ws)[0] for r in range(ws)]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def combine_nap_stats(col_sum: torch.Tensor, n_local: int, group=None) -> Tuple[torch.Tensor, int]:
    """Pass 1 of the NAP fit across ranks: (sum over all rows, total row count)."""
    rank, ws = world(group)
    cnt = torch.tensor([float(n_local)], dtype=torch.float64, device=col_sum.device)
    if ws > 1:
        dist.all_reduce(col_sum, group=group)
        dist.all_reduce(cnt, group=group)
    return col_sum, int(cnt.item())


def combine_gram(gram: torch.Tensor, group=None) -> torch.Tensor:
    """Pass 2: the centred Gram matrices of the shards add up (all shards centred with the global mean)."""
    if world(group)[1] > 1:
        dist.all_reduce(gram, group=group)
    return gram


def score_sharded(model, x_all, start_layer_index: int = 0, end_layer_index: Optional[int] = None, nap: bool = False,
                  group=None):
    """Each rank scores its ``shard_range`` of ``x_all`` (host array/tensor visible to every rank) and the
    per-sample scores are gathered: {'base','sap'[, 'nap']} of length len(x_all) on every rank."""
    from .reconstruction_aggregation import get_scores
    rank, ws = world(group)
    n = len(x_all)
    lo, hi = shard_range(n, rank, ws)
    sc = get_scores(x_all[lo:hi], model, start_layer_index, end_layer_index, nap=nap)
    return {k: gather_rows(v, n, group) for k, v in sc.items()}


def data_parallel_step(model, optimizer, x_local, group=None):
    """models/auto_encoder.py:57-77 on N ranks: every rank runs the fused step on its micro-batch with
    global BatchNorm statistics, the flat gradient is all-reduced (SUM), every rank applies the same update.
    Returns the global summed loss (float)."""
    from . import train as T
    st = T.train_state(model)
    if st.world != world(group)[1] or st.group is not group:
        T.set_data_parallel(model, group)
    model.train()
    optimizer.zero_grad()
    x = x_local.cuda(next(model.parameters()).device) if not x_local.is_cuda else x_local
    loss = model.get_loss_value(x.view(x.size(0), -1), None)
    loss.backward()
    T.allreduce_gradients(model)
    optimizer.step()
    total = loss.detach().clone()
    if st.world > 1:
        dist.all_reduce(total, group=group)
    return float(total)
