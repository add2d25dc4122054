"""CPU time of the host layout (awb_layout_build) for one bench-size window."""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import emul_lib as el  # noqa: E402
from argweaver_b200 import sim  # noqa: E402
from argweaver_b200.problem import make_problem  # noqa: E402

for internal in (False, True):
    d = sim.simulate_problem(50, 1000000, ntimes=20, seed=1000, internal=internal)
    d.pop("mappings", None)
    p, keep = make_problem(d)
    L = el.lib()
    for i in range(3):
        t = time.perf_counter()
        rc = L.emul_layout_only(C.byref(p), 1)
        print("internal=%d layout %.1f ms rc=%d" % (internal, (time.perf_counter() - t) * 1e3, rc))
