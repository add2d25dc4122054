import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from argweaver_b200 import api, sim
import oracle_lib as ol
d = sim.simulate_problem(8, 2000, ntimes=20, seed=1, internal=False)
o = ol.run_oracle(d)
b = api.Batch([d], keep_debug=True)
b.upload().setup().forward().sync()
bs = np.concatenate([[0], np.cumsum(d["blocklens"])])
blk = 17
S1 = o["nstates"][blk]
lo = o["fw_off"][blk] + (1493 - bs[blk]) * S1
print("oracle col 1493 entries:", {i: o["fw"][lo + i] for i in (33, 90, 91, 92, 93, 94)})
sws = b.debug("sw_start"); swc = b.debug("sw_cnt"); src = b.debug("sw_src"); prob = b.debug("sw_prob")
