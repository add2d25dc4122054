"""Wall-clock breakdown of one end-to-end call (host buffers -> paths) at bench size."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
from argweaver_b200 import api, sim  # noqa: E402


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    print("free/total GB:", [x / 1e9 for x in torch.cuda.mem_get_info()], flush=True)
    sites = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    keep, problems, rands = [], [], []
    for w in range(W):
        d = sim.simulate_problem(50, sites, ntimes=20, seed=1000 + w, internal=(w % 2 == 1))
        d.pop("mappings", None)
        for key in ("seqs", "ptrees", "ages", "sprs", "blocklens"):
            d[key], t = pinned(d[key])
            keep.append(t)
        r, t = pinned(np.random.RandomState(w).randint(0, 2**31 - 1, sites).astype(np.int32))
        keep.append(t)
        problems.append(d)
        rands.append(r)
    pbuf = []
    for w in range(W):
        pb, t = pinned(np.zeros(sites, np.int32))
        keep.append(t)
        pbuf.append(pb)
    ctx = api.Context(0)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = api.Batch(problems, ctx, checkpoint=True)
        t1 = time.perf_counter()
        b.upload().sync()
        t2 = time.perf_counter()
        b.setup().forward().traceback(rands).sync()
        t3 = time.perf_counter()
        paths = [b.path(i, out=pbuf[i]) for i in range(W)]
        t4 = time.perf_counter()
        tm = b.timings()
        b.close()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        print("serial    e2e %.0f ms: create %.0f upload %.0f run %.0f paths %.0f close %.0f | %s"
              % ((t5 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3,
                 (t4 - t3) * 1e3, (t5 - t4) * 1e3, tm), flush=True)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = api.Batch(problems, ctx, checkpoint=True)
        t1 = time.perf_counter()
        b.upload().setup().forward().traceback(rands)
        t2 = time.perf_counter()
        b.sync()
        t3 = time.perf_counter()
        paths = [b.path(i, out=pbuf[i]) for i in range(W)]
        t4 = time.perf_counter()
        tm = b.timings()
        b.close()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        print("pipelined e2e %.0f ms: create %.0f enqueue %.0f wait %.0f paths %.0f close %.0f | %s"
              % ((t5 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3,
                 (t4 - t3) * 1e3, (t5 - t4) * 1e3, tm), flush=True)


def b_t(b):
    return ""


if __name__ == "__main__":
    main()
