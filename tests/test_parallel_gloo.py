"""Frame sharding + reductions over 2 ranks on CPU (gloo): the host logic of the N > 1 path.

The reference splits frames with np.array_split into contiguous blocks and merges with
ndarray_sum / ndarray_vstack (analysis/base.py:566-588, analysis/results.py:185-321);
`mdanalysis_b200.parallel` does the same with collectives.  The analysis used here is a
host-only stand-in with the BlockAnalysis interface, so no CUDA call is made.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _ToyAnalysis:
    """Histogram of |x| of atom 0 per frame ('sum' result) and per-frame mean radius ('concat')."""

    def __init__(self, block):
        from mdanalysis_b200.analysis.base import Results
        self._block = block
        self.results = Results()
        self.n_frames = block.n_frames

    def _prepare(self, blk):
        self.results.count = np.zeros(8, dtype=np.float64)
        self.results.series = np.zeros((blk.n_frames, 2), dtype=np.float64)

    def run(self):
        blk = self._block
        self._prepare(blk)
        r = np.linalg.norm(blk.positions.astype(np.float64), axis=2)        # [F, n]
        self.results.count = np.histogram(r, bins=8, range=(0.0, 2.0))[0].astype(np.float64)
        self.results.series = np.stack([blk.frames.astype(np.float64), r.mean(axis=1)], axis=1)
        return self


def _worker(rank, world, port, n_frames, out):
    import torch.distributed as dist
    from mdanalysis_b200 import FrameBlock, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        pos = rng.random((n_frames, 5, 3)).astype(np.float32)
        blk = FrameBlock(pos, np.array([3, 3, 3, 90, 90, 90], dtype=np.float32))
        run = parallel.ShardedRun(_ToyAnalysis, {"count": "sum", "series": "concat"}, device="cpu")
        res = run.run(blk)
        lo, hi = parallel.shard_bounds(n_frames, world, rank)
        if rank == 0:
            np.savez(out, count=res["count"], series=res["series"], lo=lo, hi=hi)
        else:
            np.savez(out + f".r{rank}.npz", lo=lo, hi=hi)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 2, 1])
def test_sharded_run_two_ranks(tmp_path, n_frames):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, out), nprocs=2, join=True)
    got = np.load(out)
    rng = np.random.default_rng(7)
    pos = rng.random((n_frames, 5, 3)).astype(np.float32)
    r = np.linalg.norm(pos.astype(np.float64), axis=2)
    want_count = np.histogram(r, bins=8, range=(0.0, 2.0))[0]
    assert np.array_equal(got["count"], want_count)                      # ndarray_sum
    assert np.array_equal(got["series"][:, 0], np.arange(n_frames))      # ndarray_vstack: frame order kept
    assert np.allclose(got["series"][:, 1], r.mean(axis=1), rtol=0, atol=0)
    other = np.load(out + ".r1.npz")
    # contiguous blocks with the reference's remainder rule (np.array_split)
    split = np.array_split(np.arange(n_frames), 2)
    assert (int(got["lo"]), int(got["hi"])) == ((int(split[0][0]), int(split[0][-1]) + 1) if len(split[0]) else (0, 0))
    assert (int(other["lo"]), int(other["hi"])) == ((int(split[1][0]), int(split[1][-1]) + 1) if len(split[1]) else (0, 0))


class _RefToy(_ToyAnalysis):
    """Like RMSD(reference=None): measures against frame 0 of ``_refblock`` (default: its own block)."""

    def __init__(self, block):
        super().__init__(block)
        self._refblock = block

    def run(self):
        blk = self._block
        self._prepare(blk)
        ref = self._refblock.positions[0].astype(np.float64)
        d = np.sqrt(((blk.positions.astype(np.float64) - ref[None]) ** 2).sum(axis=(1, 2)))
        self.results.series = np.stack([blk.frames.astype(np.float64), d], axis=1)
        return self


class _FailsOnRank1(_ToyAnalysis):
    def run(self):
        import torch.distributed as dist
        if dist.get_rank() == 1:
            raise ValueError("bad shard")
        return super().run()


def _worker_ref(rank, world, port, out):
    import torch.distributed as dist
    from mdanalysis_b200 import FrameBlock, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        blk = FrameBlock(rng.random((6, 4, 3)).astype(np.float32))
        res = parallel.ShardedRun(_RefToy, {"series": "concat"}, device="cpu").run(blk)
        msg = ""
        try:
            parallel.ShardedRun(_FailsOnRank1, {"count": "sum"}, device="cpu").run(blk)
        except Exception as e:                                   # every rank must get here, none may hang
            msg = f"{type(e).__name__}: {e}"
        np.savez(out + f".r{rank}.npz", series=res["series"], msg=msg)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_sharded_reference_is_global_and_errors_reach_every_rank(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "ref")
    mp.spawn(_worker_ref, args=(2, _free_port(), out), nprocs=2, join=True)
    rng = np.random.default_rng(11)
    pos = rng.random((6, 4, 3)).astype(np.float64).astype(np.float32).astype(np.float64)
    want = np.sqrt(((pos - pos[0][None]) ** 2).sum(axis=(1, 2)))
    for r in range(2):
        got = np.load(out + f".r{r}.npz")
        assert np.allclose(got["series"][:, 1], want, rtol=0, atol=1e-12)   # frame 0 of the WHOLE run, on both ranks
    assert "bad shard" in str(np.load(out + ".r1.npz")["msg"])
    assert "another rank failed" in str(np.load(out + ".r0.npz")["msg"])


@pytest.mark.parametrize("n,world", [(10, 3), (98, 8), (5, 8), (1000, 8), (0, 2)])
def test_shard_frames_matches_reference_split(n, world):
    from mdanalysis_b200 import parallel
    blocks = [parallel.shard_frames(n, world, r) for r in range(world)]
    ref = np.array_split(np.arange(n), world)
    for b, r in zip(blocks, ref):
        assert np.array_equal(b, r)
    assert np.array_equal(np.concatenate(blocks), np.arange(n))


def test_collectives_are_identity_without_process_group():
    import torch
    from mdanalysis_b200 import parallel
    t = torch.arange(5, dtype=torch.float64)
    assert not parallel.is_distributed()
    assert torch.equal(parallel.allreduce_sum(t.clone()), t)
    assert torch.equal(parallel.allgather_series(t.clone(), 5), t)
