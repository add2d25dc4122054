"""awb_batch_kernel_times: per-kernel device times (CUDA events around every
launch) and the algorithmic bytes of the forward launches -- what bench.py's
roofline figure is computed from."""

import numpy as np
import pytest

from argweaver_b200 import api, sim

pytestmark = pytest.mark.gpu


def test_whole_table_bytes_and_times():
    ds = [sim.simulate_problem(10, 4000, seed=90 + i, internal=bool(i & 1)) for i in range(3)]
    rs = [np.random.RandomState(i).randint(0, 2**31 - 1, 4000).astype(np.int32)
          for i in range(3)]
    b = api.Batch(ds)
    b.upload().kernel_times(True)
    b.setup().forward().traceback(rs).sync()
    kt = b.get_kernel_times()
    # 8 B per site*state, plus nothing else: one forward launch
    assert kt["forward_launches"] == 1
    assert kt["forward_bytes"] == 8.0 * sum(b.fw(i).size for i in range(3))
    for k in ("block_setup", "switch_setup", "emit", "forward", "traceback"):
        assert kt[k] > 0.0, k
    # a read drains the counters
    kt2 = b.get_kernel_times()
    assert kt2["forward_launches"] == 0 and kt2["forward"] == 0.0
    b.close()


def test_checkpointed_counts_the_rebuilt_segments(monkeypatch):
    monkeypatch.setenv("AWB_SEG_DOUBLES", "30000")
    monkeypatch.setenv("AWB_RESIDENT_SEGS", "3")
    d = sim.simulate_problem(10, 6000, seed=95)
    r = np.random.RandomState(5).randint(0, 2**31 - 1, 6000).astype(np.int32)
    b = api.Batch([d], checkpoint=True)
    b.upload().kernel_times(True)
    b.setup().forward().sync()
    first = b.get_kernel_times()
    nseg, nres = b.segments()
    assert nseg > 4 and nres == 3
    assert first["forward_launches"] == nseg
    b.traceback([r]).sync()
    second = b.get_kernel_times()
    # the second pass recomputes everything but the resident segments: fewer
    # bytes than the first pass, in fewer launches (groups of segments)
    assert 0 < second["forward_bytes"] < first["forward_bytes"]
    assert second["forward_launches"] <= nseg
    b.close()
