import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from argweaver_b200 import api, sim
import oracle_lib as ol
k, n, T, internal, seed, site = [int(x) for x in sys.argv[1:7]]
d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=bool(internal))
o = ol.run_oracle(d)
b = api.Batch([d], keep_debug=True)
b.upload().setup().forward().sync()
fw = b.fw()
ns = np.maximum(o["nstates"], 1)
bs = np.concatenate([[0], np.cumsum(d["blocklens"])])
blk = int(np.searchsorted(bs, site, side="right") - 1)
print("blocks around:", [(int(x), int(d["blocklens"][x]), int(ns[x])) for x in range(max(0, blk - 3), min(len(ns), blk + 3))])
for s in (site - 1, site):
    bl = int(np.searchsorted(bs, s, side="right") - 1)
    S1 = ns[bl]
    lo = o["fw_off"][bl] + (s - bs[bl]) * S1
    a, r = fw[lo:lo + S1], o["fw"][lo:lo + S1]
    rel = np.abs(a - r) / np.maximum(np.abs(r), 1e-300)
    idx = np.argsort(-rel)[:6]
    print("site", s, "block", bl, "S", S1)
    st = o["states"][o["state_off"][bl]:o["state_off"][bl] + o["nstates"][bl]]
    for i in idx:
        print("   state", i, st[i] if len(st) else None, "gpu %.6e ref %.6e rel %.2e" % (a[i], r[i], rel[i]))
sws = b.debug("sw_start"); swc = b.debug("sw_cnt"); src = b.debug("sw_src"); prob = b.debug("sw_prob")
lay = b.layout()
