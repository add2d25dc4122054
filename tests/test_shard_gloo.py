"""The N>1 path on CPU: two `gloo` ranks shard a list of windows, each runs its
own slice (here through the CPU oracle -- no GPU in this container) and rank 0
gathers the per-window log-likelihoods.  This is exactly the plumbing bench.py
uses under torchrun with `nccl` (argweaver_b200/shard.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, HERE

from argweaver_b200 import shard


def test_window_plan_covers_everything():
    for n in (0, 1, 7, 8, 33):
        for world in (1, 2, 3, 8):
            plan = shard.window_plan(n, world)
            ids = [w for r in plan for w in r]
            assert ids == list(range(n))
            sizes = [len(r) for r in plan]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.window_plan(4, 0)


def test_single_process_helpers():
    assert shard.all_max(3.5) == 3.5
    assert shard.all_sum(2.0) == 2.0
    assert np.array_equal(shard.gather_window_values([1.0, 2.0], 2), [1.0, 2.0])
    with pytest.raises(ValueError):
        shard.gather_window_values([1.0], 2)


def _window_problem(w):
    from argweaver_b200 import sim
    return sim.simulate_problem(5, 150, 8, seed=300 + w, internal=(w % 2 == 1))


def _window_logz(w):
    import oracle_lib
    return oracle_lib.run_oracle(_window_problem(w))["logZ"]


def _worker(rank, world, port, nwindows, outdir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port,
                            rank=rank, world_size=world)
    try:
        mine = shard.my_windows(nwindows, rank, world)
        logz = [_window_logz(w) for w in mine]
        t = 10.0 + rank                      # stand-in for a per-rank step time
        tmax = shard.all_max(t, dist)
        total = shard.all_sum(float(len(mine)), dist)
        allz = shard.gather_window_values(logz, nwindows, dist)
        dist.barrier()
        if rank == 0:
            np.savez(os.path.join(outdir, "gathered.npz"), logz=allz, tmax=tmax,
                     total=total)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("nwindows", [4, 5])
def test_two_ranks_gather_window_loglikelihoods(tmp_path, nwindows):
    import oracle_lib
    oracle_lib.build()
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nwindows, str(tmp_path)),
             nprocs=world, join=True)
    z = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    serial = np.array([_window_logz(w) for w in range(nwindows)])
    assert np.array_equal(z["logz"], serial)          # same code, same inputs
    assert float(z["tmax"]) == 11.0
    assert float(z["total"]) == nwindows
