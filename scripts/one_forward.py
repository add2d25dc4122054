"""Run setup+forward(+traceback) once on a small problem (for ncu captures)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from argweaver_b200 import api, sim
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 1
T = int(sys.argv[4]) if len(sys.argv) > 4 else 20
ds = [sim.simulate_problem(k, n, ntimes=T, seed=100 + c) for c in range(C)]
rs = [np.random.RandomState(c).randint(0, 2**31 - 1, n).astype(np.int32) for c in range(C)]
b = api.Batch(ds)
b.upload().setup().forward().traceback(rs).sync()
print(b.timings(), "maxS", max(b.nstates(0)))
