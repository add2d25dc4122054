"""Window sharding across GPUs.

The threading HMM has no cross-window data dependency: arg-sample-genome
(reference bin/arg-sample-genome) cuts a genome into independent windows and
runs one arg-sample per window.  Here one process drives one GPU, every rank
takes a contiguous slice of the window list, and the ONLY exchange is the
gather of per-window scalars (log-likelihood, counters) on rank 0 -- there is
no collective inside the recursion.

The helpers work with any torch.distributed backend: `nccl` on the GPUs,
`gloo` in the CPU tests (tests/test_shard_gloo.py).
"""
import numpy as np


def window_plan(nwindows, world):
    """Contiguous, balanced assignment: list of window-id ranges per rank."""
    if world <= 0:
        raise ValueError("world size must be positive")
    base, extra = divmod(nwindows, world)
    plan, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        plan.append(range(start, start + n))
        start += n
    return plan


def my_windows(nwindows, rank, world):
    return window_plan(nwindows, world)[rank]


def _device(dist):
    import torch
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def all_max(x, dist=None):
    """max over ranks of a Python float (timings are reported as the slowest rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(x, dist=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_window_values(local_values, nwindows, dist=None):
    """Per-window float64 values of every rank, in global window order.

    local_values: the values of my_windows(nwindows, rank, world), in order.
    Ranks may hold different counts (nwindows not divisible by world): slices are
    padded to the longest one for the all_gather and trimmed again.
    """
    local = np.asarray(local_values, np.float64).reshape(-1)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        if len(local) != nwindows:
            raise ValueError("expected %d values, got %d" % (nwindows, len(local)))
        return local
    import torch
    world = dist.get_world_size()
    rank = dist.get_rank()
    plan = window_plan(nwindows, world)
    if len(local) != len(plan[rank]):
        raise ValueError("rank %d holds %d windows but passed %d values"
                         % (rank, len(plan[rank]), len(local)))
    width = max(len(r) for r in plan)
    dev = _device(dist)
    buf = torch.full((max(width, 1),), float("nan"), dtype=torch.float64, device=dev)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.empty(nwindows, np.float64)
    for r, rng in enumerate(plan):
        out[rng.start:rng.stop] = parts[r][:len(rng)].cpu().numpy()
    return out
