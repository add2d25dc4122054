#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the threading-HMM path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path (per-block setup, variant-site emissions,
forward recursion, stochastic traceback) over a batch of independent genome
windows resident on the GPU.  The workload is BASELINE.json configs[2]
("k=50, L=10 Mb, ntimes=20: leaf plus internal-branch (subtree) threading"),
at site compression c=10 (10^6 compressed sites per window), with
`--windows` windows per GPU (half threaded as a new leaf = external mode, half
as a re-threaded subtree = internal mode).  Multi-GPU runs give every rank its
own windows (weak scaling; windows are independent, there is no collective in
the data path -- only a small all_gather of per-window logZ at the end).

metric `value`  = sum over windows of (sum_blocks blocklen*nstates) / device time
`e2e`           = same, through the public API with host (pinned) inputs: batch
                  creation (host layout), H2D of trees/sequences/draws, all
                  kernels, D2H of the sampled paths, every step -- a stream of
                  batches: the host-only creation of batch n+1 overlaps the
                  device work of batch n (`single_batch_latency_ms` is one
                  batch alone, nothing overlapped).
`roofline`      = forward kernel: 8 B per site*state (the FP64 forward-table
                  store) / its CUDA-event time, against
                  the measured HBM copy bandwidth in MEASURED_PEAKS.json.
`cpu_baseline`  = the UNMODIFIED reference's own forward+traceback
                  (oracle/_ref/ref_bench, 1 core) on a bounded sample of the
                  same workload.

--impl reference times the reference's CPU implementation (oracle/_ref/ref_bench)
with one single-threaded worker per host core on bounded samples of the same
workload.
"""

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hmm_sites_x_states_per_sec"
UNIT = "sites*states/s"
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--sites", type=int, default=1000000,
                    help="compressed sites per window (L/c)")
    ap.add_argument("--ntimes", type=int, default=20)
    ap.add_argument("--windows", type=int, default=0,
                    help="independent windows per GPU (default: 148 with the "
                         "checkpointed table, 38 with the whole table)")
    ap.add_argument("--checkpoint", type=int, default=1,
                    help="1: AWB_CHECKPOINT (forward table kept one segment at a "
                         "time, forward recursion run twice); 0: whole table")
    ap.add_argument("--cpu-sample-sites", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.windows <= 0:
        a.windows = 148 if a.checkpoint else 38
    return a


def workload_name(a):
    return ("config3: arg-sim k=%d, L=%.0f Mb, ntimes=%d, maxtime=200e3, c=10 "
            "(%d compressed sites/window); leaf + subtree threading"
            % (a.k, a.sites * 10 / 1e6, a.ntimes, a.sites))


# --------------------------------------------------------------------- clocks

class ClockSampler(object):
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,"
              "clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smmax)) if smmax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ reference

def truncated_problem(a, nsites, seed=1, internal=False):
    from argweaver_b200 import sim
    return sim.simulate_problem(a.k, nsites, ntimes=a.ntimes, seed=seed,
                                internal=internal)


def run_ref_bench(problem, threads=1, reps=1):
    """Time the unmodified reference on a problem; returns dict or None."""
    from argweaver_b200.flatfile import write_awf
    if not os.path.exists(REF_BENCH):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "p.awf")
        write_awf(fn, problem)
        out = subprocess.run([REF_BENCH, "--in", fn, "--reps", str(reps),
                              "--threads", str(threads)], capture_output=True,
                             text=True)
    if out.returncode != 0:
        return None
    kv = dict(re.findall(r"(\w+)=([-0-9.e+]+)", out.stdout))
    return {k: float(v) for k, v in kv.items()}


def run_oracle_port(problem):
    """Fallback CPU timing with the C oracle (kind 'port')."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    n = int(np.sum(problem["blocklens"]))
    r = np.random.RandomState(1).randint(0, 2**31 - 1, n).astype(np.int32)
    t0 = time.time()
    run = ol.OracleRun(problem).setup().forward()
    t1 = time.time()
    run.traceback(r)
    t2 = time.time()
    o = run.outputs()
    ss = float(np.sum(o["nstates"].astype(np.float64) * problem["blocklens"]))
    return {"forward_s": t1 - t0, "trace_s": t2 - t1, "wall_s": t2 - t0,
            "states_sites": ss, "threads": 1}


def cpu_baseline(a):
    prob = truncated_problem(a, a.cpu_sample_sites, seed=1)
    res = run_ref_bench(prob, threads=1, reps=1)
    kind = "reference"
    if res is None:
        res = run_oracle_port(prob)
        kind = "port"
    t = res["forward_s"] + res["trace_s"]
    return {"value": res["states_sites"] / t, "unit": UNIT, "cores": 1,
            "kind": kind,
            "forward_s": res["forward_s"], "trace_s": res["trace_s"],
            "sample": "forward+traceback of one external-mode window, first "
                      "%d of %d compressed sites (k=%d, ntimes=%d), 1 core"
                      % (a.cpu_sample_sites, a.sites, a.k, a.ntimes)}


def bench_reference(a, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nsites = min(a.cpu_sample_sites, a.sites)
    prob = truncated_problem(a, nsites, seed=1)
    kind = "reference" if os.path.exists(REF_BENCH) else "port"
    times, ss = [], 0.0
    for step in range(a.warmup + a.steps):
        t0 = time.time()
        if kind == "reference":
            res = run_ref_bench(prob, threads=cores, reps=1)
            wall = res["wall_s"]
            ss = res["states_sites"] * cores
        else:
            res = run_oracle_port(prob)
            wall = res["wall_s"]
            ss = res["states_sites"]
            cores = 1
        if step >= a.warmup:
            times.append(wall)
        del t0
    t = float(np.mean(times))
    value = ss / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "k": a.k, "ntimes": a.ntimes,
                   "sites_per_window": a.sites, "compress": 10},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": kind,
                         "sample": "each step: %d concurrent single-threaded "
                                   "workers (one per host core), each running "
                                   "forward+traceback on the first %d sites of "
                                   "an external-mode window" % (cores, nsites)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ b200

def pinned_like(arr):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t.numpy(), t


def bench_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from argweaver_b200 import api, shard, sim

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic windows for this rank (pinned host memory)
    W = a.windows
    problems, rands, keep = [], [], []
    for w in range(W):
        seed = 1000 + shard.my_windows(W * world, rank, world)[w]
        d = sim.simulate_problem(a.k, a.sites, ntimes=a.ntimes, seed=seed,
                                 internal=(w % 2 == 1))
        # the simulator's SPRs keep node indices stable, i.e. the default node
        # mapping (identity except the broken node): not passed, made on device
        d.pop("mappings", None)
        for key in ("seqs", "ptrees", "ages", "sprs", "blocklens"):
            d[key], t = pinned_like(np.ascontiguousarray(d[key]))
            keep.append(t)
        r = np.random.RandomState(seed).randint(0, 2**31 - 1, a.sites)
        r, t = pinned_like(r.astype(np.int32))
        keep.append(t)
        problems.append(d)
        rands.append(r)

    ctx = api.Context(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident measurement (`value`)
    batch = api.Batch(problems, ctx, checkpoint=bool(a.checkpoint))
    batch.upload().upload_rand(rands).sync()
    ss_local = batch.total_states_sites()
    fw_bytes = sum(8.0 * batch.states_sites(i) for i in range(W))

    def step():
        batch.setup().forward().traceback()

    for _ in range(a.warmup):
        step()
    batch.sync()
    launches0 = batch.kernel_launches()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.record(0)
    stage = {"setup_ms": 0.0, "forward_ms": 0.0, "traceback_ms": 0.0}
    for _ in range(a.steps):
        step()
        if a.steps <= 8:
            tm = batch.timings()       # syncs the stream; per-stage CUDA events
            for k2 in stage:
                stage[k2] += tm[k2] / a.steps
    ctx.record(1)
    barrier()
    clocks = sampler.stop()
    ms_local = ctx.elapsed_ms(0, 1) / a.steps
    launches = (batch.kernel_launches() - launches0) // max(a.steps, 1)
    if a.steps > 8:
        stage = batch.timings()
    logz_local = [batch.logz(i) for i in range(W)]
    status = [batch.status(i) for i in range(W)]
    assert all(s == -1 for s in status), "forward hit a non-positive column"
    nseg, nres = batch.segments()
    h2d = batch.h2d_bytes() + sum(r.nbytes for r in rands)
    d2h = sum(4 * batch.nsites(i) for i in range(W))
    batch.close()

    # ---- end to end through the public API, host buffers in pinned memory
    e2e_ms_local = None
    e2e_latency_ms = None
    if not a.no_e2e:
        path_bufs = []
        for w in range(W):
            pb, t = pinned_like(np.zeros(a.sites, np.int32))
            keep.append(t)
            path_bufs.append(pb)

        # A stream of batches through api.sample_thread_stream, the way a
        # genome-wide run would drive the library: awb_batch_create is host-only
        # work (the layout), so the NEXT batch is created while the current one
        # runs on the device; every step still pays its own layout,
        # host->device copies of all its inputs, kernels, and device->host
        # copies of all its paths.
        def create():
            return api.Batch(problems, ctx, checkpoint=bool(a.checkpoint))

        nsteps = max(1, min(a.steps, 3))
        stream = api.sample_thread_stream(
            ((problems, rands) for _ in range(nsteps + 2)), ctx,
            checkpoint=bool(a.checkpoint), out=lambda i: path_bufs[i])
        next(stream)                       # warm-up batch (and batch 1 is created)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):            # (each of them also creates its successor)
            next(stream)
        torch.cuda.synchronize()
        e2e_ms_local = (time.perf_counter() - t0) * 1e3 / nsteps
        stream.close()
        # latency of ONE isolated batch (nothing overlapped): create -> paths
        barrier()
        t0 = time.perf_counter()
        b1 = create()
        b1.upload().setup().forward().traceback(rands).sync()
        _ = [b1.path(i, out=path_bufs[i]) for i in range(W)]
        b1.close()
        e2e_latency_ms = (time.perf_counter() - t0) * 1e3

    # ---- reduce over ranks: max time, sum work (argweaver_b200/shard.py)
    d_ = dist if world > 1 else None
    ms = shard.all_max(ms_local, d_)
    ss = shard.all_sum(ss_local, d_)
    fwd_ms = shard.all_max(stage["forward_ms"], d_)
    e2e_ms = shard.all_max(e2e_ms_local, d_) if e2e_ms_local is not None else None
    # the only exchange of the workload: per-window log-likelihoods to rank 0
    logz_all = shard.gather_window_values(logz_local, W * world, d_)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else \
            "fallback 6650 GB/s (B200_PROFILING.md)"
        achieved = fw_bytes / (fwd_ms * 1e-3) / 1e9
        # DRAM traffic of the same kernel from the committed ncu capture
        # (profiles/), valid only for the configuration it was taken on
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles",
                                             "r1_forward_traffic.json")))
            c = tj["config"]
            if (c["k"], c["ntimes"], c["sites_per_window"], c["windows_per_gpu"],
                    c.get("checkpoint", 0)) == (a.k, a.ntimes, a.sites, W,
                                                a.checkpoint):
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": ss / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(a), "k": a.k, "ntimes": a.ntimes,
                "sites_per_window": a.sites, "compress": 10,
                "windows_per_gpu": W, "modes": "even windows external (new "
                "leaf), odd windows internal (subtree)",
                "l2": "inputs larger than L2: %.1f GB of forward table per GPU "
                      "streamed per step" % (fw_bytes / 1e9),
                "parallelism": "windows sharded across GPUs, one CTA per window",
                "table": ("checkpointed: %d segments of <= 128 MiB per window, "
                          "%d segment tables kept per window (all the device "
                          "memory allows); the traceback rebuilds the other "
                          "%d segments (forward recursion runs %.2f times per "
                          "step)" % (nseg, nres, nseg - nres,
                                     1.0 + (nseg - nres) / float(nseg))
                          if a.checkpoint else "whole forward table resident"),
            },
            "clocks": clocks,
            "e2e": None if e2e_ms is None else {
                "value": ss / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms,
                "single_batch_latency_ms": e2e_latency_ms,
                "note": "stream of batches through the public API, host (pinned) "
                        "buffers: the host-only layout of batch n+1 "
                        "(awb_batch_create) overlaps the device work of batch n; "
                        "all copies of every batch are inside the timed region"},
            "gpu_launches": int(launches),
            "stage_ms": stage,
            # SURVEY 8d also asks for the forward-only figure (per-segment
            # emission + forward kernels of the first pass)
            "forward_only": {"value": ss / (fwd_ms * 1e-3) if fwd_ms > 0 else None,
                             "unit": UNIT},
            "roofline": {
                "kernel": "awb_forward_fast_kernel", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": fw_bytes,
                "note": "8 B per site*state (FP64 forward-table store) over the "
                        "forward stage of the step (per-segment emission + forward "
                        "kernels); with the checkpointed table the traceback stage "
                        "runs the same kernels once more"},
            "logz_mean": float(np.mean(logz_all)),
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        bench_reference(a, rank, world)
        return
    bench_b200(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
