"""Multi-GPU drivers: one process per GPU (torch.distributed, NCCL over NVLink); the terrain path row-shards.

The reference's analogue is the overlap-tiling of ``geoutils.map_overlap_multiproc_save`` (terrain.py:412-466): tiles
overlap by ``depth`` = 1 (3x3 fits/windows), 2 (Florinsky / 5x5).  Here every rank owns a contiguous block of rows and
receives ``depth`` halo rows from each neighbour (grouped NCCL send/recv, a few hundred KB -- latency bound); the first
and last shard see the raster border as NaN through the kernel's out-of-bounds fill.  Invariant (tested): the sharded
result equals the single-GPU result bit-for-bit.
"""

from __future__ import annotations

from typing import Any, Sequence

import torch
import torch.distributed as dist

from . import _engine


def halo_depth(surface_attributes: Sequence[str], windowed_indexes: Sequence[str], surface_fit: str,
               window_size: int) -> int:
    """Overlap rule of terrain.py:417-432."""
    d = 0
    if windowed_indexes:
        d = window_size // 2
    if surface_attributes:
        d = max(d, 2 if surface_fit.lower() == "florinsky" else 1)
    return d


def _peer(group: Any, group_rank: int) -> int:
    """Global rank of a group-local rank (dist.P2POp addresses peers by GLOBAL rank; for the default group the two
    coincide)."""
    if group is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


def _all_ranks_ok(ok: bool, group: Any, device: torch.device) -> bool:
    """Collective validity check: every rank learns whether ANY rank failed a precondition, so that all of them raise
    instead of the healthy ones hanging in the next collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ok
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class RowShard:
    """Halo-row exchange for a row-sharded raster.  ``buf`` is (rows + 2*depth, cols): [top halo | core | bottom halo].
    ``rank`` / ``world`` are ranks WITHIN ``group`` (north to south)."""

    def __init__(self, rank: int, world: int, depth: int, group: Any = None) -> None:
        self.rank, self.world, self.depth, self.group = rank, world, depth, group

    @property
    def has_top(self) -> bool:
        return self.rank > 0

    @property
    def has_bottom(self) -> bool:
        return self.rank < self.world - 1

    def start_exchange(self, buf: torch.Tensor, rows: int) -> list[Any]:
        """Post the grouped send/recv of the halo rows and return the work handles (the transfers run on the backend's
        own stream; ``finish_exchange`` makes the current stream wait for them)."""
        d = self.depth
        if self.world == 1 or d == 0:
            return []
        ops = []
        if self.has_top:
            up = _peer(self.group, self.rank - 1)
            ops.append(dist.P2POp(dist.isend, buf[d:2 * d], up, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[0:d], up, self.group))
        if self.has_bottom:
            down = _peer(self.group, self.rank + 1)
            ops.append(dist.P2POp(dist.isend, buf[rows:rows + d], down, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[rows + d:rows + 2 * d], down, self.group))
        return dist.batch_isend_irecv(ops)

    @staticmethod
    def finish_exchange(works: list[Any]) -> None:
        for w in works:
            w.wait()

    def exchange(self, buf: torch.Tensor, rows: int) -> None:
        self.finish_exchange(self.start_exchange(buf, rows))

    def view(self, buf: torch.Tensor, rows: int) -> tuple[int, int, torch.Tensor]:
        """(row_begin, row_end, view) to hand to the kernel: the view drops the halo block on a raster border so that
        those rows are out-of-bounds (= NaN) for the stencil."""
        d = self.depth
        top = 0 if self.has_top else d
        bottom = rows + 2 * d if self.has_bottom else rows + d
        r_begin = d - top
        return r_begin, r_begin + rows, buf[top:bottom]

    def prepare(self, buf: torch.Tensor, rows: int) -> tuple[int, int, torch.Tensor]:
        """Exchange halos (blocking the stream), then return ``view(buf, rows)``."""
        self.exchange(buf, rows)
        return self.view(buf, rows)

    def run_overlapped(self, buf: torch.Tensor, rows: int, launch: Any) -> None:
        """Halo exchange overlapped with compute: ``launch(view, row_begin, row_end, out_row0)`` is called for the
        interior rows (which only read this shard's own rows) while the halo rows are in flight, then -- once they have
        arrived -- for the ``depth``-row strips next to each neighbour.  ``out_row0`` is the first output row of the
        piece relative to the shard.  Shards too thin to have an interior fall back to exchange-then-compute."""
        d = self.depth
        r0, r1, view = self.view(buf, rows)
        if self.world == 1 or d == 0 or rows <= 2 * d:
            self.exchange(buf, rows)
            launch(view, r0, r1, 0)
            return
        works = self.start_exchange(buf, rows)
        lo = r0 + (d if self.has_top else 0)
        hi = r1 - (d if self.has_bottom else 0)
        launch(view, lo, hi, lo - r0)
        self.finish_exchange(works)
        if self.has_top:
            launch(view, r0, lo, 0)
        if self.has_bottom:
            launch(view, hi, r1, hi - r0)


def sharded_terrain_attribute(local_rows: torch.Tensor, resolution: float, surface_attributes: Sequence[str] = (),
                              windowed_indexes: Sequence[str] = (), surface_fit: str = "Florinsky",
                              window_size: int = 3, group: Any = None, out: torch.Tensor | None = None,
                              **kwargs: Any) -> torch.Tensor:
    """Each rank passes its contiguous block of rows (CUDA tensor, rank order within ``group`` = north to south) and
    gets the (n_attr, rows, cols) planes of its block; halo rows travel over NCCL while the interior rows are already
    being computed."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    d = halo_depth(surface_attributes, windowed_indexes, surface_fit, window_size)
    rows, cols = local_rows.shape
    if not _all_ranks_ok(world == 1 or rows >= d, group, local_rows.device):
        raise ValueError(f"each shard needs at least {d} rows")
    buf = torch.empty((rows + 2 * d, cols), dtype=local_rows.dtype, device=local_rows.device)
    buf[d:d + rows] = local_rows
    n_attr = len(surface_attributes) + len(windowed_indexes)
    if out is None:
        out = torch.empty((n_attr, rows, cols), dtype=local_rows.dtype, device=local_rows.device)

    def launch(view: torch.Tensor, rb: int, re: int, out_row0: int) -> None:
        _engine.terrain_fused(view, resolution, surface_attributes, windowed_indexes, surface_fit=surface_fit,
                              window_size=window_size, row_begin=rb, row_end=re,
                              out=out[:, out_row0:out_row0 + (re - rb)], **kwargs)

    RowShard(rank, world, d, group).run_overlapped(buf, rows, launch)
    return out


def _exchange_rows(core: torch.Tensor, depth: int, rank: int, world: int, group: Any = None
                   ) -> tuple[torch.Tensor | None, torch.Tensor | None]:
    """(rows of the shard above, rows of the shard below) -- `depth` boundary rows from each neighbour, None at a raster
    border."""
    rows, cols = core.shape
    buf = torch.empty((rows + 2 * depth, cols), dtype=core.dtype, device=core.device)
    buf[depth:depth + rows] = core
    RowShard(rank, world, depth, group).exchange(buf, rows)
    top = buf[:depth].clone() if rank > 0 else None
    bot = buf[rows + depth:].clone() if rank < world - 1 else None
    return top, bot


def sharded_nuth_kaab(ref_rows: torch.Tensor, tba_rows: torch.Tensor, inlier_rows: torch.Tensor | None = None,
                      transform: Any = None, max_iterations: int = 10, tolerance: float = 0.001, bin_sizes: int = 72,
                      halo: int = 16, group: Any = None) -> tuple[tuple[float, float, float], int]:
    """Row-sharded Nuth & Kaab fit (BASELINE config 5): every rank passes its contiguous block of rows of the reference
    and to-be-aligned DEMs (CUDA float32, rank order = north to south).  One halo row of the reference (np.gradient)
    and `halo` rows of the to-be-aligned DEM (bilinear gather under the running shift; |row shift| must stay below
    halo-1 pixels) are exchanged once over NCCL; per iteration only the radix-select histograms, the moments and the
    aspect range are all-reduced (<= 72 x 256 x 8 B), and every rank runs the same 72-point curve_fit.  Returns the same
    (easting, northing, vertical) offsets on every rank and the global number of valid pixels."""
    import scipy.optimize

    from . import coreg

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if ref_rows.shape != tba_rows.shape or ref_rows.dim() != 2:
        raise ValueError("reference and to-be-aligned shards must be 2-D and of the same shape")
    if not _all_ranks_ok(world == 1 or ref_rows.shape[0] >= halo, group, ref_rows.device):
        raise ValueError(f"each shard needs at least {halo} rows")
    ref_rows = ref_rows.to(torch.float32).contiguous()
    tba_rows = tba_rows.to(torch.float32).contiguous()
    if world > 1:
        ref_halo = _exchange_rows(ref_rows, 1, rank, world, group)
        tba_halo = _exchange_rows(tba_rows, halo, rank, world, group)
    else:
        ref_halo, tba_halo = (None, None), (None, None)
    state = coreg._NKState(ref_rows, tba_rows, inlier_rows, group=group, ref_halo=ref_halo, tba_halo=tba_halo)
    state.sharded = world > 1
    n_valid = state._n_valid_t.clone()
    if world > 1:
        dist.all_reduce(n_valid, group=group)
    if int(n_valid.item()) == 0:
        raise ValueError("There is no valid points common to the input and auxiliary data.")
    offsets = coreg._iterate_nuth_kaab(state, transform, int(bin_sizes), scipy.optimize.curve_fit, tolerance,
                                       max_iterations)
    return offsets, int(n_valid.item())
