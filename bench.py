#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the threading-HMM path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config 1|2|3|4|5]

A *step* is one pass of the hot path (per-block setup, variant-site emissions,
forward recursion, stochastic traceback) over a batch of independent genome
windows resident on the GPU.  The default workload is BASELINE.json configs[2]
("k=50, L=10 Mb, ntimes=20: leaf plus internal-branch (subtree) threading") at
site compression c=10 (10^6 compressed sites per window), half of the windows
threaded as a new leaf (external mode), half as a re-threaded subtree (internal
mode).  `--config` selects the other named shapes (CONFIGS below).  Multi-GPU
runs give every rank its own windows (weak scaling; windows are independent,
there is no collective in the data path -- only a small all_gather of
per-window logZ at the end); config 5 is the one strong-scaling case (8 windows
in total, shared out over the ranks).

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
                  same workload, one external-mode and one internal-mode window.
`parity`        = the same sample problems run on the GPU in the same process
                  and compared with what the reference just produced: sampled
                  paths (identical draws) and forward rows.

--impl reference times the reference's CPU implementation (oracle/_ref/ref_bench)
with one single-threaded worker per host core on bounded samples of the same
workload (half of the workers on the external-mode, half on the internal-mode
sample).
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

# BASELINE.json configs[i-1].  `windows`: independent windows per GPU (0: one per
# SM with the checkpointed table; 592 = four per SM for the small shapes, whose
# forward kernel is a 6-warp CTA; 48 = what fits for config 4's 2.8 GB of
# per-block tables per window); `total_windows`: fixed number shared out over
# the ranks (strong scaling).
CONFIGS = {
    1: dict(k=8, sites=10000, ntimes=20, windows=592, checkpoint=0,
            name="config1: arg-sim -k 8 -L 100000 -N 10000 -r 1.6e-8 -m 1.8e-8, "
                 "--ntimes 20 --maxtime 200e3 -c 10 (README quick start)"),
    2: dict(k=20, sites=100000, ntimes=20, windows=592, checkpoint=0,
            name="config2: k=20, L=1 Mb, ntimes=20, c=10: full-thread resampling"),
    3: dict(k=50, sites=1000000, ntimes=20, windows=148, checkpoint=1,
            name="config3: arg-sim k=50, L=10 Mb, ntimes=20, maxtime=200e3, c=10"),
    4: dict(k=100, sites=1000000, ntimes=40, windows=48, checkpoint=1,
            name="config4: k=100, L=10 Mb, ntimes=40, c=10: large state space"),
    5: dict(k=50, sites=1000000, ntimes=20, total_windows=8, checkpoint=0,
            name="config5: arg-sample-genome, k=50, 8 independent 10 Mb windows "
                 "sharded across the GPUs"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--sites", type=int, default=0,
                    help="compressed sites per window (L/c)")
    ap.add_argument("--ntimes", type=int, default=0)
    ap.add_argument("--windows", type=int, default=0,
                    help="independent windows per GPU (default: per config)")
    ap.add_argument("--checkpoint", type=int, default=-1,
                    help="1: AWB_CHECKPOINT (forward table kept one segment at a "
                         "time, forward recursion run twice); 0: whole table")
    ap.add_argument("--cpu-sample-sites", type=int, default=200000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dense-seqs", action="store_true",
                    help="upload dense character rows instead of the variant "
                         "columns only (the .sites form)")
    ap.add_argument("--mcmc-iters", type=int, default=-1,
                    help="iterations of the arg-sample pair (metric ii); default: "
                         "100 for config 1, 3 for config 2, none otherwise; 0: skip")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.k = a.k or c["k"]
    a.sites = a.sites or c["sites"]
    a.ntimes = a.ntimes or c["ntimes"]
    if a.checkpoint < 0:
        a.checkpoint = c["checkpoint"]
    a.strong = "total_windows" in c and a.windows <= 0
    a.total_windows = c.get("total_windows", 0)
    if a.windows <= 0:
        a.windows = c.get("windows", 0)
    if a.windows <= 0 and not a.strong:
        a.windows = 148 if a.checkpoint else 38
    return a


def config_dict(a):
    """The workload, identical for both arms (the driver compares them)."""
    return {"workload": workload_name(a), "k": a.k, "ntimes": a.ntimes,
            "sites_per_window": a.sites, "compress": 10,
            "modes": "even windows external (new leaf), odd windows internal "
                     "(subtree)",
            "l2": "inputs larger than L2: every step streams its windows' forward "
                  "tables (8 B per site*state, GBs per GPU; `run.table_gb_per_gpu` "
                  "in the b200 arm) and per-block tables through HBM"}


def workload_name(a):
    return ("%s (%d compressed sites/window); leaf + subtree threading"
            % (CONFIGS[a.config]["name"], a.sites))


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

def sample_problems(a):
    """The bounded CPU sample of the workload: the first `cpu_sample_sites`
    sites of one external-mode and one internal-mode window."""
    from argweaver_b200 import sim
    n = min(a.cpu_sample_sites, a.sites)
    return [sim.simulate_problem(a.k, n, ntimes=a.ntimes, seed=1 + i,
                                 internal=bool(i)) for i in range(2)], n


def ref_bench_start(fn, threads=1, reps=1, out=None, seed=1, stride=1):
    cmd = [REF_BENCH, "--in", fn, "--reps", str(reps), "--threads", str(threads),
           "--rand-seed", str(seed)]
    if out:
        cmd += ["--out", out, "--fw-stride", str(stride)]
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                            text=True)


def ref_bench_result(proc):
    so, se = proc.communicate()
    if proc.returncode != 0:
        raise RuntimeError("ref_bench failed: " + se[-1000:])
    kv = dict(re.findall(r"(\w+)=([-0-9.e+]+)", so))
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
            "states_sites": ss, "threads": 1, "path": o["path"]}


def gpu_parity(problems, seeds, refs, ctx):
    """Run the sample problems on the GPU (whole table) and compare with what
    the reference produced for them: paths (same libc rand() draws) and the
    forward rows the reference dumped."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_lib
    from argweaver_b200 import api
    libc = ctypes.CDLL("libc.so.6")
    out = {"path_identical": True, "first_divergence": None, "fw_max_rel": 0.0,
           "sites_compared": 0, "fw_rows_compared": 0,
           "against": "oracle/_ref/ref_bench (unmodified reference), same "
                      "problems, same srand() seeds"}
    for d, seed, ref in zip(problems, seeds, refs):
        n = int(np.sum(d["blocklens"]))
        libc.srand(seed)
        r = np.array([libc.rand() for _ in range(n)], np.int32)
        b = api.Batch([d], ctx)
        b.upload().setup().forward().traceback([r]).sync()
        path = b.path(0)
        bad = np.nonzero(path != ref["path"])[0]
        if bad.size:
            out["path_identical"] = False
            if out["first_divergence"] is None:
                out["first_divergence"] = int(bad.max())   # the walk runs backwards
        lay = b.layout(0)
        mine = ref_lib.rows_of(b.fw(0), lay["fw_off"], ref["nstates"],
                               d["blocklens"], ref["fw_sites"])
        with np.errstate(all="ignore"):
            rel = np.abs(mine - ref["fw"]) / np.maximum(
                np.maximum(np.abs(mine), np.abs(ref["fw"])), 1e-300)
        rel[mine == ref["fw"]] = 0.0
        out["fw_max_rel"] = max(out["fw_max_rel"], float(np.max(rel)))
        out["sites_compared"] += n
        out["fw_rows_compared"] += int(len(ref["fw_sites"]))
        b.close()
    return out


def cpu_baseline(a, ctx=None):
    """1 core of the unmodified reference on the bounded sample; with `ctx`,
    the GPU runs the same problems and the results are compared (parity)."""
    from argweaver_b200.flatfile import read_awf, write_awf
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    problems, n = sample_problems(a)
    parity = None
    if not os.path.exists(REF_BENCH):
        res = [run_oracle_port(p) for p in problems]
        kind = "port"
    else:
        import ref_lib
        kind = "reference"
        res, refs = [], []
        with tempfile.TemporaryDirectory() as tmp:
            for i, p in enumerate(problems):
                fn = os.path.join(tmp, "p%d.awf" % i)
                fo = os.path.join(tmp, "r%d.awf" % i)
                write_awf(fn, ref_lib.problem_for_file(p))
                # enough repetitions for ~5 s of CPU work per mode
                est = a.k * 6.0 * n / 3.0e7
                reps = int(max(1, min(200, round(5.0 / max(est, 1e-3)))))
                res.append(ref_bench_result(ref_bench_start(
                    fn, 1, reps, out=fo, seed=11 + i, stride=max(1, n // 400))))
                refs.append(read_awf(fo))
        if ctx is not None:
            parity = gpu_parity(problems, [11, 12], refs, ctx)
    t = sum(r["forward_s"] + r["trace_s"] for r in res)
    ss = sum(r["states_sites"] for r in res)
    out = {"value": ss / t, "unit": UNIT, "cores": 1, "kind": kind,
           "forward_s": sum(r["forward_s"] for r in res),
           "trace_s": sum(r["trace_s"] for r in res),
           "sample": "forward+traceback of one external-mode and one "
                     "internal-mode window, first %d of %d compressed sites "
                     "(k=%d, ntimes=%d), 1 core" % (n, a.sites, a.k, a.ntimes)}
    return out, parity


ARG_SAMPLE = os.path.join(ROOT, "oracle", "_ref", "arg-sample")
ARG_SAMPLE_B200 = os.path.join(ROOT, "oracle", "_ref", "arg-sample-b200")


def mcmc_pair(a, iters):
    """BASELINE metric (ii), arg-sample MCMC iterations per second: the
    UNMODIFIED reference binary (oracle/_ref/arg-sample, 1 core -- it is
    single-threaded) and the same binary with its threading path on the device
    (oracle/_ref/arg-sample-b200 = the reference's objects + the adapter
    oracle/sample_thread_b200.cpp over libargweaver_b200.so), same generated
    .sites file, same seed.  One chain issues its thread samples one after
    another, so this is the single-window latency of the path, not its batched
    throughput.  iters/s = iterations / sum of the per-iteration `sample time`
    lines of the run's log (arg-sample.cpp:676)."""
    from argweaver_b200 import sim
    if not (os.path.exists(ARG_SAMPLE) and os.path.exists(ARG_SAMPLE_B200)):
        return {"unavailable": "oracle/_ref/arg-sample[-b200] not built"}
    out = {"metric": "arg_sample_mcmc_iters_per_sec", "iters": iters,
           "config": "k=%d, %d compressed sites (-c 10), ntimes=%d, -x 1"
                     % (a.k, a.sites, a.ntimes)}
    with tempfile.TemporaryDirectory() as tmp:
        seqs = sim.simulate_arg(a.k - 1, a.sites, a.ntimes, seed=77)[8]
        sites = os.path.join(tmp, "gen.sites")
        sim.write_sites(sites, seqs, compress=10)
        stats = {}
        for name, binary in (("reference", ARG_SAMPLE), ("b200", ARG_SAMPLE_B200)):
            o = os.path.join(tmp, name)
            cmd = [binary, "-s", sites, "-N", "10000", "-r", "1.6e-8", "-m",
                   "1.8e-8", "--ntimes", str(a.ntimes), "--maxtime", "200e3",
                   "-c", "10", "-n", str(iters), "-x", "1", "-q", "-o", o]
            env = dict(os.environ)
            env["AWB_ADAPTER_REPORT"] = "1"
            t0 = time.time()
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            wall = time.time() - t0
            if r.returncode != 0:
                return {"unavailable": "%s failed: %s" % (name, r.stderr[-300:])}
            tms = []
            for m in re.finditer(r"^sample time:\s*([0-9.]+)\s*(us|ms|s|m|h)\b",
                                 open(o + ".log").read(), re.M):
                tms.append(float(m.group(1)) * {"us": 1e-6, "ms": 1e-3, "s": 1.0,
                                                "m": 60.0, "h": 3600.0}[m.group(2)])
            stats[name] = open(o + ".stats").read()
            out[name] = {"iters_per_s": len(tms) / sum(tms) if tms else None,
                         "wall_s": wall, "cores": 1}
            if name == "b200":
                m = re.search(r"device thread samples: (\d+) reference fallbacks: "
                              r"(\d+) device seconds: ([0-9.]+)", r.stderr)
                if m:
                    out[name].update(device_thread_samples=int(m.group(1)),
                                     reference_fallbacks=int(m.group(2)),
                                     device_seconds=float(m.group(3)))
        out["stats_identical"] = stats["reference"] == stats["b200"]
        if out["reference"]["iters_per_s"] and out["b200"]["iters_per_s"]:
            out["speedup"] = out["b200"]["iters_per_s"] / out["reference"]["iters_per_s"]
    return out


def bench_reference(a, rank, world):
    if rank != 0:
        return
    from argweaver_b200.flatfile import write_awf
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cores = os.cpu_count() or 1
    problems, nsites = sample_problems(a)
    kind = "reference" if os.path.exists(REF_BENCH) else "port"
    times, ss = [], 0.0
    with tempfile.TemporaryDirectory() as tmp:
        fns = []
        if kind == "reference":
            import ref_lib
            for i, p in enumerate(problems):
                fns.append(os.path.join(tmp, "p%d.awf" % i))
                write_awf(fns[-1], ref_lib.problem_for_file(p))
        # small configurations: repeat inside the workers so that a step is
        # seconds of work, not process start-up
        est = a.k * 6.0 * nsites / 3.0e7
        reps = int(max(1, min(500, round(1.0 / max(est, 1e-3)))))
        for step in range(a.warmup + a.steps):
            if kind == "reference":
                # half of the cores on the external-mode sample, half on the
                # internal-mode one, all at once
                th = [cores - cores // 2, cores // 2]
                t0 = time.time()
                procs = [ref_bench_start(fns[i], th[i], reps) for i in range(2)
                         if th[i] > 0]
                res = [ref_bench_result(p) for p in procs]
                wall = time.time() - t0
                ss = sum(r["states_sites"] * r["threads"] * r["reps"] for r in res)
            else:
                res = [run_oracle_port(p) for p in problems]
                wall = sum(r["wall_s"] for r in res)
                ss = sum(r["states_sites"] for r in res)
                cores = 1
            if step >= a.warmup:
                times.append(wall)
    t = float(np.mean(times))
    value = ss / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong" if a.strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                         "kind": kind,
                         "sample": "each step: %d concurrent single-threaded "
                                   "workers (one per host core; half on an "
                                   "external-mode, half on an internal-mode "
                                   "window), each running forward+traceback %d "
                                   "time(s) on the first %d sites of its window; "
                                   "wall time includes process start and reading "
                                   "the problem file" % (cores, reps, nsites)},
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
    if a.strong:
        mine = list(shard.my_windows(a.total_windows, rank, world))
        total_windows = a.total_windows
    else:
        mine = list(shard.my_windows(a.windows * world, rank, world))
        total_windows = a.windows * world
    W = len(mine)
    problems, rands, keep = [], [], []
    for w in mine:
        seed = 1000 + w
        d = sim.simulate_problem(a.k, a.sites, ntimes=a.ntimes, seed=seed,
                                 internal=(w % 2 == 1))
        # the simulator's SPRs keep node indices stable, i.e. the default node
        # mapping (identity except the broken node): not passed, made on device
        d.pop("mappings", None)
        if not a.dense_seqs:
            # the alignment as its variant columns (what a .sites file holds)
            d = sim.pack_problem(d)
        for key in ("seqs", "var_pos", "var_cols", "ptrees", "ages", "sprs",
                    "blocklens"):
            if key not in d:
                continue
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

    ms_local = 0.0
    ss_local = 0.0
    fw_bytes = 0.0
    stage = {"setup_ms": 0.0, "forward_ms": 0.0, "traceback_ms": 0.0}
    launches = 0
    logz_local = []
    nseg, nres = 1, 1
    h2d = d2h = 0
    clocks = None
    fast_path = None
    e2e_ms_local = None
    e2e_latency_ms = None
    if W > 0:
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
    ktimes = None
    if W > 0:
        batch.kernel_times(True)   # CUDA events around every launch, on its stream
        ctx.record(0)
        for _ in range(a.steps):
            step()
            if a.steps <= 8:
                tm = batch.timings()       # syncs the stream; per-stage CUDA events
                for k2 in stage:
                    stage[k2] += tm[k2] / a.steps
        ctx.record(1)
    barrier()
    clocks = sampler.stop()
    if W > 0:
        ms_local = ctx.elapsed_ms(0, 1) / a.steps
        ktimes = batch.get_kernel_times()
        launches = (batch.kernel_launches() - launches0) // max(a.steps, 1)
        if a.steps > 8:
            stage = batch.timings()
        logz_local = [batch.logz(i) for i in range(W)]
        status = [batch.status(i) for i in range(W)]
        assert all(s == -1 for s in status), "forward hit a non-positive column"
        nseg, nres = batch.segments()
        fast_path = batch.fast_path()
        h2d = batch.h2d_bytes() + sum(r.nbytes for r in rands)
        d2h = sum(4 * batch.nsites(i) for i in range(W))
        batch.close()

    # ---- end to end through the public API, host buffers in pinned memory
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

        nsteps = max(1, a.steps)
        if W > 0:
            stream = api.sample_thread_stream(
                ((problems, rands) for _ in range(nsteps + 2)), ctx,
                checkpoint=bool(a.checkpoint), out=lambda i: path_bufs[i])
            next(stream)                   # warm-up batch (and batch 1 is created)
        barrier()
        t0 = time.perf_counter()
        if W > 0:
            for _ in range(nsteps):        # (each of them also creates its successor)
                next(stream)
        torch.cuda.synchronize()
        e2e_ms_local = (time.perf_counter() - t0) * 1e3 / nsteps
        if W > 0:
            stream.close()
        # latency of ONE isolated batch (nothing overlapped): create -> paths
        barrier()
        if W > 0:
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
    fw_bytes_all = shard.all_sum(fw_bytes, d_)
    # the forward kernel's own launches (both passes of a checkpointed table):
    # slowest rank's kernel time, bytes summed over the ranks
    k4_ms = shard.all_max(ktimes["forward"] / a.steps if ktimes else 0.0, d_)
    k4_bytes_all = shard.all_sum(ktimes["forward_bytes"] / a.steps if ktimes else 0.0, d_)
    e2e_ms = shard.all_max(e2e_ms_local, d_) if e2e_ms_local is not None else None
    h2d_all = shard.all_sum(float(h2d), d_)
    d2h_all = shard.all_sum(float(d2h), d_)
    # the only exchange of the workload: per-window log-likelihoods to rank 0
    logz_all = shard.gather_window_values(logz_local, total_windows, d_)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else \
            "fallback 6650 GB/s (B200_PROFILING.md)"
        # per GPU: the forward kernel's launches of one step (timed with CUDA
        # events around each launch on its stream) over the bytes they computed
        achieved = (k4_bytes_all / world) / (k4_ms * 1e-3) / 1e9 if k4_ms > 0 else 0.0
        k4_n = ktimes["forward_launches"] // max(a.steps, 1) if ktimes else 0
        # DRAM traffic of the same kernel: NOT measured in this run -- taken
        # from the committed ncu capture of the same configuration, if any
        traffic, traffic_src = None, None
        for fn in ("r2_forward_traffic.json", "r1_forward_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
                c = tj["config"]
                if (c["k"], c["ntimes"], c["sites_per_window"], c["windows_per_gpu"],
                        c.get("checkpoint", 0)) == (a.k, a.ntimes, a.sites, W,
                                                    a.checkpoint):
                    traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                    traffic_src = ("profiles/%s (ncu --set full capture of this "
                                   "configuration, not measured in this run)" % fn)
                    break
            except Exception:
                pass
        line = {
            "metric": METRIC, "value": ss / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if a.strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a),
            "run": {
                "windows_per_gpu": W if not a.strong else
                "%d windows in total over %d GPU(s)" % (total_windows, world),
                "table_gb_per_gpu": fw_bytes / 1e9,
                "parallelism": "windows sharded across GPUs, one CTA per window",
                "forward_kernel": fast_path,
                "table": ("checkpointed: %d segments of <= 64 MiB per window, "
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
                "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "ms_per_step": e2e_ms, "steps": max(1, a.steps),
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
                "kernel": "awb_forward_fast_kernel" if fast_path == "fast"
                else "awb_forward_kernel", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "launches_per_step": k4_n,
                "kernel_ms_per_step": k4_ms,
                "algorithmic_bytes_per_step": k4_bytes_all / world,
                "algorithmic_bytes_per_launch": (k4_bytes_all / world / k4_n) if k4_n else None,
                "note": "8 B per site*state (FP64 forward-table store) that the "
                        "forward kernel's launches of one step computed (first pass "
                        "and, with the checkpointed table, the segments the second "
                        "pass rebuilds), over the summed duration of those "
                        "launches, CUDA events around each launch"},
            "kernel_ms": None if not ktimes else {
                k2: ktimes[k2] / a.steps for k2 in api.Batch.KERNEL_CLASSES},
            "logz_mean": float(np.mean(logz_all)),
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"], line["parity"] = cpu_baseline(a, ctx)
            iters = a.mcmc_iters
            if iters < 0:
                iters = {1: 100, 2: 3}.get(a.config, 0)
            if iters > 0:
                ctx.close()            # the sampler process takes the device
                ctx = None
                line["mcmc"] = mcmc_pair(a, iters)
        print(json.dumps(line), flush=True)
    if ctx is not None:
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
