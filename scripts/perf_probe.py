"""Quick device-time probe of the pipeline stages (not the contract bench)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from argweaver_b200 import api, sim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--sites", type=int, default=100000)
    ap.add_argument("--ntimes", type=int, default=20)
    ap.add_argument("--chains", type=str, default="1")
    ap.add_argument("--internal", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--checkpoint", type=int, default=0)
    ap.add_argument("--packed", type=int, default=0,
                    help="alignment as variant columns (the .sites form)")
    ap.add_argument("--rho", type=float, default=1.6e-8,
                    help="recombination rate of the simulation (lower: longer blocks)")
    ap.add_argument("--ktimes", type=int, default=0,
                    help="also print the per-kernel times (CUDA events around each launch)")
    a = ap.parse_args()
    for C in [int(x) for x in a.chains.split(",")]:
        t0 = time.time()
        ds = [sim.simulate_problem(a.k, a.sites, ntimes=a.ntimes, seed=100 + c,
                                   internal=bool(a.internal), rho=a.rho) for c in range(C)]
        if a.packed:
            ds = [sim.pack_problem(d) for d in ds]
        print("blocks per window: %d (%.1f sites per block)"
              % (len(ds[0]["blocklens"]), a.sites / float(len(ds[0]["blocklens"]))))
        rs = [np.random.RandomState(c).randint(0, 2**31 - 1, a.sites).astype(np.int32)
              for c in range(C)]
        tg = time.time() - t0
        t0 = time.time()
        b = api.Batch(ds, checkpoint=bool(a.checkpoint))
        tc = time.time() - t0
        for rep in range(a.reps):
            t0 = time.time()
            b.upload()
            b.sync()
            tu = time.time() - t0
            if a.ktimes:
                b.kernel_times(True)
            b.setup().forward().traceback(rs).sync()
            if a.ktimes:
                kt = b.get_kernel_times()
                print("kernel ms: " + " ".join("%s %.2f" % (k, v) for k, v in kt.items()
                                                if k != "forward_bytes"), flush=True)
            tm = b.timings()
            ss = b.total_states_sites()
            print("k=%d n=%d T=%d C=%d int=%d | gen %.1fs create %.3fs upload %.3fs | "
                  "setup %.2f ms fwd %.2f ms tb %.2f ms | fwd %.3e st-sites/s "
                  "(%.3f us/site/chain) all %.3e | fwd HBM %.1f GB/s"
                  % (a.k, a.sites, a.ntimes, C, a.internal, tg, tc, tu,
                     tm["setup_ms"], tm["forward_ms"], tm["traceback_ms"],
                     ss / tm["forward_ms"] * 1e3,
                     tm["forward_ms"] * 1e3 / a.sites,
                     ss / (tm["setup_ms"] + tm["forward_ms"] + tm["traceback_ms"]) * 1e3,
                     8 * ss / tm["forward_ms"] * 1e3 / 1e9), flush=True)
        b.close()


if __name__ == "__main__":
    main()
