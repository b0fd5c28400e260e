"""Frame sharding across GPUs and the reductions that follow.

The reference parallelises over frames only: ``AnalysisBase._setup_computation_groups``
splits the frame indices with ``np.array_split`` into contiguous blocks
(``analysis/base.py:566-588``), workers run ``_compute`` on their block, and
``ResultsGroup.merge`` reduces on the host -- ``ndarray_sum`` for histograms,
``ndarray_vstack`` for per-frame series (``analysis/results.py:185-321``).

Here: one process per GPU (torchrun), rank r owns block r of the same split,
and the reductions are collectives on device tensors -- NCCL over NVLink on
GPUs, gloo on CPU for the tests:

  histogram counts / volume_cum   all_reduce(SUM)     (ndarray_sum)
  rmsd / contact series           all_gather (padded) (ndarray_vstack / hstack)

Only these few hundred bytes to kilobytes cross the links; coordinates never do.
"""
from __future__ import annotations

import numpy as np


def shard_frames(n_frames: int, world_size: int, rank: int) -> np.ndarray:
    """Frame indices of ``rank``: block ``rank`` of ``np.array_split(arange(n), world)``."""
    return np.array_split(np.arange(n_frames), world_size)[rank]


def shard_bounds(n_frames: int, world_size: int, rank: int):
    idx = shard_frames(n_frames, world_size, rank)
    if len(idx) == 0:
        return 0, 0
    return int(idx[0]), int(idx[-1]) + 1


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed() -> bool:
    try:
        dist = _dist()
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


def allreduce_sum(t):
    """In-place SUM over ranks of a torch tensor (device for NCCL, CPU for gloo)."""
    if is_distributed():
        _dist().all_reduce(t, op=_dist().ReduceOp.SUM)
    return t


def allgather_series(local, n_total: int, world_size: int | None = None):
    """Concatenate per-rank series (first dim = this rank's frames) in rank order.

    Ranks may own different numbers of frames (array_split remainder rule), so
    every rank pads to the largest block before the collective.
    """
    import torch
    if not is_distributed():
        return local
    dist = _dist()
    world = world_size or dist.get_world_size()
    sizes = [len(np.array_split(np.arange(n_total), world)[r]) for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


class ShardedRun:
    """Run a block analysis on this rank's frames and reduce like ResultsGroup.merge.

    ``make_analysis(block)`` builds the analysis for a FrameBlock; ``reduce`` maps
    result names to 'sum' (histograms) or 'concat' (per-frame series).
    """

    def __init__(self, make_analysis, reduce, device=None):
        self.make_analysis = make_analysis
        self.reduce = dict(reduce)
        self.device = device

    def run(self, block, world_size=None, rank=None):
        import torch
        if world_size is None:
            world_size = _dist().get_world_size() if is_distributed() else 1
        if rank is None:
            rank = _dist().get_rank() if is_distributed() else 0
        lo, hi = shard_bounds(block.n_frames, world_size, rank)
        dev = self.device or ("cuda" if (torch.cuda.is_available() and is_distributed()
                                         and _dist().get_backend() == "nccl") else "cpu")
        # A rank that fails locally still enters the collectives (the others would wait in
        # all_reduce / all_gather for ever): the error is agreed on first, then raised everywhere.
        error = None
        ana = None
        try:
            if hi > lo:
                local = block.slice_frames(slice(lo, hi))
                ana = self.make_analysis(local)
                # "reference = the trajectory itself" means frame ref_frame of the WHOLE run
                # (analysis/rms.py:560-567), not of this rank's shard
                if getattr(ana, "_refblock", None) is local:
                    ana._refblock = block
                ana.run()
                shapes = ana.results
            else:
                shapes = self._result_shapes(block)
        except Exception as e:                              # noqa: BLE001 -- re-raised below, on every rank
            error = e
        flag = torch.tensor([0.0 if error is None else 1.0], dtype=torch.float64).to(dev)
        if allreduce_sum(flag).item() > 0:
            if error is not None:
                raise error
            raise RuntimeError("another rank failed in its shard of the run")
        out = {}
        for name, how in self.reduce.items():
            val = np.asarray(shapes[name], dtype=np.float64)
            if how == "sum":
                if hi <= lo:
                    val = np.zeros_like(val)
                t = torch.as_tensor(val).to(dev)
                out[name] = allreduce_sum(t).cpu().numpy()
            elif how == "concat":
                if hi <= lo:
                    val = val[:0]
                t = torch.as_tensor(val).to(dev)
                out[name] = allgather_series(t, block.n_frames, world_size).cpu().numpy()
            else:
                raise ValueError(how)
        self.analysis = ana
        return out

    def _result_shapes(self, block):
        """Result arrays of a rank WITHOUT frames (world_size > n_frames): the analysis is prepared
        on the first frame of the whole block -- never on an empty one, whose ``_prepare`` would
        index frame 0 -- and only the shapes / dtypes of what it allocates are used."""
        probe_block = block.slice_frames(slice(0, 1))
        probe = self.make_analysis(probe_block)
        if hasattr(probe, "_b2"):
            probe._b2_run = probe._b2
        probe.n_frames = probe_block.n_frames
        probe._prepare(probe_block)
        return probe.results
