import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from argweaver_b200 import api, sim
import oracle_lib as ol
k, n, T, internal, seed = [int(x) for x in sys.argv[1:6]]
d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=bool(internal))
o = ol.run_oracle(d)
b = api.Batch([d], keep_debug=True)
b.upload().setup().forward().sync()
fw = b.fw()
ns = np.maximum(o["nstates"], 1)
bs = np.concatenate([[0], np.cumsum(d["blocklens"])])
kind = b.debug("kind")
bad = 0
for blk in range(len(ns)):
    S1 = ns[blk]
    for i in range(bs[blk], bs[blk + 1]):
        lo = o["fw_off"][blk] + (i - bs[blk]) * S1
        a, r = fw[lo:lo + S1], o["fw"][lo:lo + S1]
        err = np.max(np.abs(a - r) / np.maximum(np.abs(r), 1e-300))
        if err > 1e-9:
            print("site", i, "block", blk, "i_in_block", i - bs[blk], "blen", d["blocklens"][blk],
                  "S", S1, "kind", kind[i], "err %.3e" % err, "nbad", int((np.abs(a - r) > 1e-9 * np.abs(r)).sum()),
                  "sum gpu %.6f" % a.sum(), "first bad idx", int(np.argmax(np.abs(a - r) / np.maximum(np.abs(r), 1e-300))))
            bad += 1
            if bad > 12:
                sys.exit(0)
print("done, bad sites:", bad, "logz", b.logz(), o["logZ"])
