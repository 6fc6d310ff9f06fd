"""World-size-2 gloo tests (CPU) of the multi-process host logic: row sharding + ordered gather, and the
NAP-fit exchange steps (column sums, centred Gram) -- two ranks that each see half of the train rows must end
with the statistics, and therefore the fit, of one process that sees all rows.  The diffs come from the oracle
(no CUDA here); on the GPU box the same combination runs over NCCL inside Engine.nap_fit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icra2021_multimodal_ad_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from icra2021_multimodal_ad_b200.engine import nap_fit_from_stats
        from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
        from oracle import rapp_oracle as RO
        torch.set_num_threads(2)
        D, btl, nl, seed, n = 64, 100, 5, 11, 301          # odd row count: ragged shards
        sd = synth_state_dict(D, btl, nl, seed)
        x, _ = synth_windows(n, D, seed + 1, anomaly_rate=0.0)
        lo, hi = P.shard_range(n, rank, world)
        d_local = torch.from_numpy(RO.concat_diffs(RO.get_diffs(x[lo:hi], sd), 0, 2)).double()
        # pass 1: column sums -> global mean
        s, total = P.combine_nap_stats(d_local.sum(0), hi - lo)
        mu = (s / total).float()
        # pass 2: Gram of the shard centred with the GLOBAL mean
        c = d_local - mu.double()
        gram = P.combine_gram(c.t() @ c)
        fit = nap_fit_from_stats(mu, gram, total, factor="eigen")
        # ordered gather of per-sample results
        local_score = (d_local.float() ** 2).mean(1)
        gathered = P.gather_rows(local_score, n)
        if rank == 0:
            d_all = torch.from_numpy(RO.concat_diffs(RO.get_diffs(x, sd), 0, 2)).double()
            mu1 = d_all.mean(0)
            c1 = d_all - mu1.float().double()
            g1 = c1.t() @ c1
            ok = dict(total=total == n,
                      mu=float((mu.double() - mu1).abs().max()) < 1e-7,
                      gram=float((gram - g1).abs().max() / g1.abs().max()) < 1e-12,
                      var=bool(torch.isfinite(fit["var"]).all()),
                      order=float((gathered - (d_all.float() ** 2).mean(1)).abs().max()) < 1e-6,
                      k=fit["vt"].shape[0] == min(n, d_all.shape[1]))
            q.put(ok)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_rows():
    for n in (0, 1, 7, 8, 1000, 10_000_000):
        for w in (1, 2, 3, 8):
            parts = [P.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_range(10, 3, 3)


def test_two_rank_nap_statistics_and_gather_match_one_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    ok = q.get(timeout=10)
    assert all(ok.values()), ok
