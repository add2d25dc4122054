#!/usr/bin/env python
"""Latency of ONE thread sample (what one arg-sample chain sees): wall time of
each API stage with a sync after it, for single windows of the named shapes."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argweaver_b200 import api, sim

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="8:2000:20,8:10000:20,20:100000:20,50:200000:20")
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
ctx = api.Context(0)
for shape in a.shapes.split(","):
    k, n, T = [int(x) for x in shape.split(":")]
    d = sim.simulate_problem(k, n, ntimes=T, seed=5)
    r = np.random.RandomState(1).randint(0, 2**31 - 1, n).astype(np.int32)
    acc = {}
    for rep in range(a.reps + 2):
        t = [time.perf_counter()]
        b = api.Batch([d], ctx); t.append(time.perf_counter())
        b.upload(); ctx.sync(); b.sync(); t.append(time.perf_counter())
        b.setup(); ctx.sync(); t.append(time.perf_counter())
        b.forward(); ctx.sync(); t.append(time.perf_counter())
        b.traceback([r]); ctx.sync(); t.append(time.perf_counter())
        b.sync(); p = b.path(0); lz = b.logz(0); t.append(time.perf_counter())
        b.close(); t.append(time.perf_counter())
        if rep >= 2:
            for nm, x0, x1 in zip(("create", "upload", "setup", "forward", "traceback",
                                   "results", "close"), t[:-1], t[1:]):
                acc[nm] = acc.get(nm, 0.0) + (x1 - x0) * 1e3 / a.reps
    # the one-shot call, everything queued at once
    t0 = time.perf_counter()
    for rep in range(a.reps):
        api.sample_thread(d, r, ctx=ctx) if hasattr(api, "sample_thread") else None
    one = (time.perf_counter() - t0) * 1e3 / a.reps
    print("k=%d n=%d T=%d: " % (k, n, T) +
          " ".join("%s %.2f" % kv for kv in acc.items()) +
          " | sum %.2f ms, one-shot %.2f ms" % (sum(acc.values()), one), flush=True)
ctx.close()
