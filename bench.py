#!/usr/bin/env python
"""Benchmark of the hot path: fused AE + RaPP anomaly scoring (base + SAP + NAP) of synthetic
1728-d multimodal windows, BASELINE.json configs[1] ("AE fp32 ... SAP/NAP scoring on 1 B200,
synthetic data of data_config.json dims").

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, libmmad.so)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A step = every rank scores ``--batch`` windows (base, SAP and NAP scores for each).  One JSON line
is printed by rank 0.  ``value`` is device-resident throughput, ``e2e`` goes through the
host-buffer C-ABI call (pinned host memory -> H2D -> scores -> D2H inside the timed region).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, BTL, NL = 1728, 100, 5          # sensor=All (utils/data_loaders.py:16-19), novelty_detection.py:147-148
W_ENC = [1728, 1402, 1076, 751, 425, 100]
DPRIME = sum(W_ENC)                # 5482
MAC_ENC = sum(a * b for a, b in zip(W_ENC[:-1], W_ENC[1:]))
FLOP_SAP = 2 * 3 * MAC_ENC         # enc(x) + dec + enc(xhat): 30 605 754 per window
FLOP_NAP_ROT = 2 * DPRIME * DPRIME  # rotation (d-mu) V: 60 104 648 per window
N_FIT = 8192                       # NAP fit set (>= D' so K = D')
DEFAULT_PRECISION = "f16f8"        # MMAD_DEFAULT_PRECISION overrides; f16x3 and fp32 are measured with --precision
DTYPE_NAME = {"fp32": "f32", "f16x3": "f16x3 split (fp32-equivalent)", "f16": "f16",
              "f16f8": "f16f8 split: fp16 hi*hi + fp8(e4m3) cross terms, fp32 accumulate (DESIGN.md section 3)"}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="mmad", choices=["mmad", "reference"])
    p.add_argument("--batch", type=int, default=4 * 148 * 128,
                   help="windows per rank per step (default 75 776 = four waves of 148 SMs x 128-row tiles)")
    p.add_argument("--precision", default=os.environ.get("MMAD_BENCH_PRECISION", "auto"))
    p.add_argument("--no-nap", action="store_true")
    p.add_argument("--cpu-sample", type=int, default=4096)
    p.add_argument("--no-extras", action="store_true", help="skip the train-step and streaming-latency sections")
    p.add_argument("--train-batch", type=int, default=256)
    return p.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 5 ms during the timed region (NVML; the region is ~100 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_ev = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_ev.wait(0.005)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------
def cpu_scoring_rate(n_sample, sd, fit, with_nap, repeats=1):
    """reconstruction_aggregation.get_diffs + utils/metric SAP (+ NAP rotate/standardise) through the
    oracle port (torch-CPU fp32, all host threads).  Returns (samples/s, cores)."""
    from icra2021_multimodal_ad_b200.utils.synth import synth_windows
    from oracle import rapp_oracle as RO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, _ = synth_windows(n_sample, D, 1236)
    nf = None
    if with_nap:
        nf = RO.NapFit.__new__(RO.NapFit)
        nf.mu, nf.v, nf.mu2, nf.var = fit["mu"].cpu(), fit["vt"].cpu().t().contiguous(), fit["mu2"].cpu(), fit["var"].cpu()

    def once():
        d = RO.get_diffs(x, sd, batch_size=256)
        RO.recon_score(d[0])
        RO.sap_score(d)
        if nf is not None:
            nf.score(RO.concat_diffs(d))
    once()   # warm-up
    t0 = time.perf_counter()
    for _ in range(repeats):
        once()
    dt = (time.perf_counter() - t0) / repeats
    return n_sample / dt, cores


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python
    reference itself cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
    from oracle import rapp_oracle as RO
    sd = synth_state_dict(D, BTL, NL, 0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = min(args.cpu_sample, 2048)
    fit = None
    if not args.no_nap:   # CPU fit with K = D' = 5482 like the GPU arm (fit time is not part of the metric)
        xtr, _ = synth_windows(5632, D, 1234, anomaly_rate=0.0)
        nf = RO.NapFit(RO.concat_diffs(RO.get_diffs(xtr, sd, batch_size=256)))
        fit = {"mu": nf.mu, "vt": nf.v.t().contiguous(), "mu2": nf.mu2, "var": nf.var}
    for _ in range(max(args.warmup, 1)):
        cpu_scoring_rate(n, sd, fit, not args.no_nap)
    t0 = time.perf_counter()
    rates = [cpu_scoring_rate(n, sd, fit, not args.no_nap)[0] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    rate = float(np.mean(rates))
    line = {"impl": "reference", "metric": "anomaly-scored samples/sec (SAP+NAP)", "value": rate, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / rate,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, os.environ.get("MMAD_DEFAULT_PRECISION", DEFAULT_PRECISION) if args.precision == "auto" else args.precision),
            "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "arithmetic": "fp32 (torch CPU)",
                             "sample": f"{n} windows per step, get_diffs(batch 256)+base+SAP" + ("" if args.no_nap else "+NAP score (K=5482 fit)")},
            "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(line))


def workload_config(args, precision):
    return {"workload": "configs[1]: fused AE + RaPP scoring (base+SAP" + ("" if args.no_nap else "+NAP") +
            ") of synthetic 1728-d windows", "input_size": D, "btl_size": BTL, "n_layers": NL,
            "windows_per_rank_per_step": args.batch, "nap_fit_rows": 0 if args.no_nap else N_FIT,
            "precision": precision, "l2": "per-step input (%.0f MB) exceeds the 126 MB L2" % (args.batch * D * 4 / 1e6),
            "parallelism": f"sample-sharded x{args.gpus}, no data-path collective"}


def bench_train(dev, local, world, batch, steps, warmup, precision, vib=False):
    """AE train samples/s (BASELINE metric, second half): AutoEncoder.step (models/auto_encoder.py:57-77) =
    train-mode forward + backward + Adam, one call per step, loss read back to the host every step like the
    reference.  Data parallel at world > 1: BatchNorm statistics and the flat gradient are all-reduced."""
    import types
    import torch.distributed as dist
    from icra2021_multimodal_ad_b200 import train as T
    from icra2021_multimodal_ad_b200.model_builder import get_model
    from icra2021_multimodal_ad_b200.models.auto_encoder import AutoEncoder
    from icra2021_multimodal_ad_b200.optim import Adam
    from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows
    cfg = argparse.Namespace(input_size=D, btl_size=BTL, n_layers=NL, gpu_id=local, precision=precision, vib=vib, beta_kl=1.0)
    model = get_model(cfg)
    model.load_state_dict(synth_state_dict(D, BTL, NL, 0, enc_out=2 * BTL if vib else None))
    opt = Adam(model.parameters(), lr=1e-3)
    if world > 1:
        T.set_data_parallel(model)
    eng = types.SimpleNamespace(model=model, optimizer=opt, config=cfg)
    xh, _ = synth_windows(batch, D, 1234 + int(os.environ.get("RANK", "0")), anomaly_rate=0.0)
    xh = xh.pin_memory()
    xd = xh.to(dev)

    def step(x):
        if world == 1:
            return AutoEncoder.step(eng, (x, None))
        model.train()
        opt.zero_grad()
        loss = model.get_loss_value(x.cuda(local), None)
        loss.backward()
        T.allreduce_gradients(model)
        opt.step()
        return (float(loss.detach()), )

    def timed(x, n):
        for _ in range(warmup):
            step(x)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step(x)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n
    ms_dev = timed(xd, steps)
    ms_e2e = timed(xh, steps)
    flop = 56366196.0 * batch * world     # fwd + dW + dX per sample
    return {"metric": "VIB-AE train samples/sec" if vib else "AE train samples/sec", "batch_per_gpu": batch, "value": world * batch / (ms_dev / 1e3),
            "ms_per_step": ms_dev, "unit": "samples/s", "algorithmic_tflops": flop / (ms_dev / 1e3) / 1e12,
            "e2e": {"value": world * batch / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": batch * D * 4,
                    "d2h_bytes_per_step": 4}, "optimizer": "mmad multi-tensor Adam", "gemm": {"fp32": "fp32 CUDA-core", "f16x3": "tcgen05 f16x3 split", "f16": "tcgen05 f16", "f16f8": "tcgen05 f16x3 split"}[precision],
            "graph": True, "collectives": None if world == 1 else
            "library-owned NCCL communicator: BatchNorm statistics all-reduced inside the captured step, flat gradient once per step"}


def bench_stream(eng, batches=(1, 8, 10, 64), calls=300, warm=50):
    """configs[4]: realtime_tester-style scoring (test_file/realtime_tester.py:291-309): one host->device->host
    call per window batch, base + SAP (nap=False there).  p50/p99 wall latency per call."""
    out = {}
    for b in batches:
        x = np.ascontiguousarray(np.random.default_rng(b).random((b, D), dtype=np.float32))
        for _ in range(warm):
            eng.score_host(x, 0, NL + 1, base=True, sap=True, nap=False)
        ts = []
        for _ in range(calls):
            t0 = time.perf_counter()
            eng.score_host(x, 0, NL + 1, base=True, sap=True, nap=False)
            ts.append(time.perf_counter() - t0)
        ts = np.sort(np.asarray(ts)) * 1e6
        out[str(b)] = {"p50_us": float(ts[len(ts) // 2]), "p99_us": float(ts[int(len(ts) * 0.99)]),
                       "p50_us_per_window": float(ts[len(ts) // 2] / b)}
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from icra2021_multimodal_ad_b200 import _lib
    from icra2021_multimodal_ad_b200.model_builder import get_model
    from icra2021_multimodal_ad_b200.utils.synth import synth_state_dict, synth_windows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    L = _lib.lib()

    precision = args.precision
    if precision == "auto":
        precision = os.environ.get("MMAD_DEFAULT_PRECISION", DEFAULT_PRECISION)
    sd = synth_state_dict(D, BTL, NL, 0)
    cfg = argparse.Namespace(input_size=D, btl_size=BTL, n_layers=NL, gpu_id=local, precision=precision)
    model = get_model(cfg).eval()
    model.load_state_dict(sd)
    eng = model.engine()

    # ---- NAP fit (setup, untimed): rows sharded over ranks, sum + Gram all-reduced ----
    fit, fit_s = None, 0.0
    if not args.no_nap:
        per = N_FIT // world
        xtr, _ = synth_windows(per, D, 1234 + rank, anomaly_rate=0.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fit = eng.nap_fit(xtr.to(dev), 0, NL + 1, distributed=world > 1)
        torch.cuda.synchronize()
        fit_s = time.perf_counter() - t0

    # ---- device-resident inputs (generated on the host from the seed, then uploaded) ----
    B = args.batch
    xh_small, _ = synth_windows(8192, D, 1236 + rank)
    x_host = xh_small.repeat((B + 8191) // 8192, 1)[:B].contiguous().pin_memory()
    x_dev = x_host.to(dev)
    want_nap = not args.no_nap

    def step_dev():
        return eng.score(x_dev, 0, NL + 1, base=True, sap=True, nap=want_nap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step_dev()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.mmad_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()        # ncu --profile-from-start off captures exactly the timed region
    ev0.record()
    for _ in range(args.steps):
        out = step_dev()
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = L.mmad_launch_count() - launches0
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-kernel timing of the dominant kernel (fused GEMM) with CUDA events, same workload ----
    import ctypes as C
    buf = (C.c_double * 3)()
    _lib.check(L.mmad_profile_begin(eng._h))
    psteps = min(args.steps, 3)
    for _ in range(psteps):
        step_dev()
    _lib.check(L.mmad_profile_end(eng._h, buf))
    gemm_ms, gemm_flops, gemm_launches = buf[0], buf[1], buf[2]
    pk, pk_kind = peaks()
    achieved_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused GEMM, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        tj = tj.get(precision, tj)            # one record per precision mode
        if tj.get("precision") == precision and tj.get("batch") == args.batch:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    # what the tensor pipe actually executes per window in this mode: 3 (f16x3) or 1 MMA per chain product, and the
    # triangular NAP factor's half of the rotation products
    # tensor work per product in fp16-pass units: f16x3 three fp16 MMAs; f16f8 one fp16 MMA + one fp8 MMA over twice the
    # contraction length at twice the rate (= one more unit)
    mma_passes = {"f16x3": 3, "f16f8": 2}.get(precision, 1)
    executed_per_window = mma_passes * (FLOP_SAP + (FLOP_NAP_ROT / 2 if want_nap else 0))
    executed_tf = value / world * executed_per_window / 1e12 if precision != "fp32" else None
    pipe_pct = None
    try:   # time-weighted tensor-pipe activity of the fused GEMM launches from the committed ncu launch list
        import csv
        rows = [r for r in csv.DictReader(l for l in open(os.path.join(ROOT, "profiles", "r1_launches_scoring_%s.csv" % precision)) if not l.startswith("=="))]
        t, a = {}, {}
        for r in rows:
            if "gemm_tc" in r["Kernel Name"]:
                v = float(r["Metric Value"].replace(",", ""))
                (t if r["Metric Name"].startswith("gpu__time") else a)[r["ID"]] = v
        if t:
            pipe_pct = sum(t[k] * a.get(k, 0.0) for k in t) / sum(t.values())
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": traffic,
                "kernel": "fused layer GEMM (%s)" % precision, "peak_source": pk_kind + " bf16 sustained",
                "launches_timed": int(gemm_launches), "gemm_share_of_step": gemm_ms / psteps / (ms / args.steps),
                "executed_mma_tflops": executed_tf, "executed_frac_of_peak": executed_tf / peak_tf if executed_tf and peak_tf else None,
                "ncu_tensor_pipe_active_pct": pipe_pct,
                "note": "achieved counts ONE product per MAC of the reference's dense algorithm; f16x3 issues 3 fp16 MMAs per product "
                        "(cap = peak/3 for dense work), f16f8 one fp16 + one double-length fp8 MMA (= 2 fp16-pass units, cap = peak/2); "
                        "the triangular NAP factor executes half of the rotation"}

    # ---- end to end through the host-buffer C-ABI call (pinned host input, scores back on host) ----
    xh_np = x_host.numpy()
    for _ in range(max(1, args.warmup // 2)):
        eng.score_host(xh_np, 0, NL + 1, base=True, sap=True, nap=want_nap)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        res = eng.score_host(xh_np, 0, NL + 1, base=True, sap=True, nap=want_nap)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = {"value": world * B * e2e_steps / float(t.item()), "unit": "samples/s",
           "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * 4 * (3 if want_nap else 2), "steps": e2e_steps}
    assert np.allclose(res["sap"][:1024], out["sap"][:1024].cpu().numpy(), rtol=1e-5)

    extras = {}
    if not args.no_extras:
        tsteps = max(3, min(args.steps * 3, 30))
        extras["train"] = bench_train(dev, local, world, args.train_batch, tsteps, 3, precision)
        extras["train_vib"] = bench_train(dev, local, world, args.train_batch, tsteps, 3, precision, vib=True)   # configs[3]
        if world == 1:
            extras["train_b7000"] = bench_train(dev, local, world, 7000, max(3, tsteps // 3), 2, precision)
            extras["stream_latency"] = bench_stream(eng)
    # ---- the same step in the full fp16 split (three fp16 MMAs per product), for comparison ----
    if precision == "f16f8" and not args.no_extras:
        eng.set_precision("f16x3")
        if want_nap:
            eng.nap_fit(xtr.to(dev), 0, NL + 1, distributed=world > 1)
        for _ in range(2):
            step_dev()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            step_dev()
        a1.record()
        barrier()
        t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extras["f16x3"] = {"value": world * B * args.steps / (float(t.item()) / 1e3), "unit": "samples/s",
                           "ms_per_step": float(t.item()) / args.steps, "dtype": DTYPE_NAME["f16x3"]}
    if rank == 0:
        cpu_rate, cores = cpu_scoring_rate(args.cpu_sample, sd, fit, want_nap) if world == 1 else (None, None)
        flop_per_window = FLOP_SAP + (FLOP_NAP_ROT if want_nap else 0)
        line = {"metric": "anomaly-scored samples/sec (SAP+NAP)", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAME[precision],
                "data": "synthetic", "config": workload_config(args, precision), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "clocks": clocks,
                "algorithmic_tflops": value * flop_per_window / 1e12, "nap_fit_s": fit_s}
        line.update(extras)
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "samples/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample} windows, oracle get_diffs(batch 256)+base+SAP" +
                                              ("+NAP score with the same fit" if want_nap else "")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
