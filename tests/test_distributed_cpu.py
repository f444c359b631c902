"""CPU: multi-process host logic with the gloo backend (world_size 2) and pure-host helpers of the variogram path."""

from __future__ import annotations

import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_worker(rank: int, world: int, port: int, depth: int, rows: int, cols: int) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xdem_b200.distributed import RowShard

    full = torch.arange(world * rows * cols, dtype=torch.float32).reshape(world * rows, cols)
    buf = torch.full((rows + 2 * depth, cols), float("nan"))
    buf[depth:depth + rows] = full[rank * rows:(rank + 1) * rows]
    shard = RowShard(rank, world, depth)
    r0, r1, view = shard.prepare(buf, rows)
    # the view must be exactly the rows [rank*rows - depth, (rank+1)*rows + depth) of the full raster, clipped
    lo = max(0, rank * rows - depth)
    hi = min(world * rows, (rank + 1) * rows + depth)
    assert torch.equal(view, full[lo:hi]), (rank, view.shape)
    assert r1 - r0 == rows and torch.equal(view[r0:r1], full[rank * rows:(rank + 1) * rows])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("depth", [1, 2])
def test_rowshard_halo_exchange_gloo(depth: int) -> None:
    port = _free_port()
    mp.spawn(_halo_worker, args=(2, port, depth, 6, 10), nprocs=2, join=True)


def _hist_worker(rank: int, world: int, port: int) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the reductions the variogram / NK paths issue: sum of int64 counts + float64 sums, min of keys
    cnt = torch.tensor([rank + 1, 10 * (rank + 1)], dtype=torch.int64)
    s = torch.tensor([0.5 * (rank + 1)], dtype=torch.float64)
    key = torch.tensor([100 - rank], dtype=torch.int64)
    dist.all_reduce(cnt)
    dist.all_reduce(s)
    dist.all_reduce(key, op=dist.ReduceOp.MIN)
    assert cnt.tolist() == [3, 30] and s.item() == 1.5 and key.item() == 99
    dist.destroy_process_group()


def _subgroup_worker(rank: int, world: int, port: int) -> None:
    """World of 3 processes; the raster is sharded over the sub-group [1, 2] (group ranks 0, 1), which does not start at
    global rank 0: halo peers must be translated to GLOBAL ranks (dist.P2POp addresses peers globally).  Also exercises
    the overlapped schedule: interior rows first, neighbour strips after the halo rows arrived."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xdem_b200.distributed import RowShard, _all_ranks_ok

    sub = dist.new_group(ranks=[1, 2])
    if rank in (1, 2):
        gr, gw, rows, cols, depth = dist.get_rank(sub), 2, 7, 5, 2
        full = torch.arange(gw * rows * cols, dtype=torch.float32).reshape(gw * rows, cols)
        buf = torch.full((rows + 2 * depth, cols), float("nan"))
        buf[depth:depth + rows] = full[gr * rows:(gr + 1) * rows]
        shard = RowShard(gr, gw, depth, sub)
        calls = []
        out = torch.full((rows, cols), float("nan"))

        def launch(view: torch.Tensor, rb: int, re: int, out_row0: int) -> None:
            calls.append((rb, re, out_row0))
            # a stand-in "stencil": every output row needs `depth` rows above and below unless at the raster border
            for r in range(rb, re):
                lo, hi = r - depth, r + depth
                rows_needed = view[max(lo, 0):min(hi, view.shape[0] - 1) + 1]
                assert not torch.isnan(rows_needed).any(), (gr, r)
                out[out_row0 + (r - rb)] = view[r]

        shard.run_overlapped(buf, rows, launch)
        assert torch.equal(out, full[gr * rows:(gr + 1) * rows])
        # interior first, then the strip next to the single neighbour
        assert len(calls) == 2 and calls[0][1] - calls[0][0] == rows - depth and calls[1][1] - calls[1][0] == depth
        # collective precondition check: one failing rank makes every rank see the failure
        assert _all_ranks_ok(True, sub, torch.device("cpu")) is True
        assert _all_ranks_ok(gr != 1, sub, torch.device("cpu")) is False
    dist.barrier()
    dist.destroy_process_group()


def test_rowshard_subgroup_not_starting_at_rank0_gloo() -> None:
    mp.spawn(_subgroup_worker, args=(3, _free_port()), nprocs=3, join=True)


def test_histogram_allreduce_gloo() -> None:
    mp.spawn(_hist_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_halo_depth_rule() -> None:
    """terrain.py:417-432."""
    from xdem_b200.distributed import halo_depth

    assert halo_depth(["slope"], [], "Horn", 3) == 1
    assert halo_depth(["slope"], [], "Florinsky", 3) == 2
    assert halo_depth([], ["roughness"], "Florinsky", 5) == 2
    assert halo_depth(["slope"], ["roughness"], "ZevenbergThorne", 5) == 2
    assert halo_depth([], ["roughness"], "Florinsky", 3) == 1


@pytest.mark.parametrize("gsd", [1.0, 5.0, 0.5, 30.0, 2.5])
def test_edge_thresholds_match_float64_semantics(gsd: float) -> None:
    """d < edge (float64, as scipy pdist / skgstat compare) <=> d2 < T for every integer squared distance."""
    from oracle import variogram_oracle as vo
    from xdem_b200.spatialstats import edge_thresholds

    maxlag = gsd * math.hypot(99, 99)
    for edges in (vo.default_bins(gsd, maxlag), list(np.linspace(0, maxlag, 21)[1:])):
        T = edge_thresholds(edges, gsd)
        d2 = np.arange(0, 2 * 99 * 99 + 2, dtype=np.int64)
        d = np.sqrt((gsd * gsd) * d2.astype(np.float64))
        for e, t in zip(edges, T):
            assert np.array_equal(d < e, d2 < t), (gsd, e, t)


def test_unit_partition_covers_every_tile_once() -> None:
    """Work units (i-group, chunk of j-groups >= i) enumerate the upper triangle exactly once."""
    chunk = 32
    for G in (1, 5, 32, 33, 100):
        i = np.arange(G, dtype=np.int64)
        prefix = np.concatenate([[0], np.cumsum((G - i + chunk - 1) // chunk)])
        seen = np.zeros((G, G), dtype=int)
        for u in range(int(prefix[-1])):
            gi = int(np.searchsorted(prefix, u, side="right") - 1)
            c = u - prefix[gi]
            for gj in range(gi + c * chunk, min(G, gi + (c + 1) * chunk)):
                seen[gi, gj] += 1
        assert np.array_equal(seen, np.triu(np.ones((G, G), dtype=int)))
