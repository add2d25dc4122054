import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from argweaver_b200 import api
g = load_golden("tests/golden/sim1_internal_uniform.npz")
b = api.Batch([g], keep_debug=True)
b.upload().setup().forward().sync()
print("status", b.status(), "logz", b.logz())
fw = b.fw()
print("fw[0:6]", fw[:6])
print("ref fw[0:6]", g["fw"][:6])
lay = b.layout(); print(lay["fw_off"][:4], b.nstates()[:4])
for nm in ("sc_row","sc_cnt","sc_start","sc_ch"):
    try:
        print(nm, b.debug(nm)[:8])
    except Exception as e:
        print(nm, "n/a", e)
print("nrm", b.debug("sink")[:12])
print("fsum", b.debug("fsum")[:19*4].reshape(4,19)[:, :4])
