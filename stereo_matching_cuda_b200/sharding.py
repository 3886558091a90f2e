"""Multi-GPU partitioning of the stereo path.  One process per GPU.

Two ways the path shards, both absent from the reference (single device, main.cu:44-48):

* batches of pairs: pair i -> rank i mod world, no communication at all;
* row strips of one large frame: every stage is row-local except the two cascaded vertical
  box passes, so a strip needs 2*radius = 18 extra INPUT rows above and below.  Those rows
  (gray/RGB image rows, never cost or a/b planes) are exchanged once with the two
  neighbouring ranks -- point-to-point send/recv over NCCL (NVLink) -- before the fused
  kernel runs with global-row window clipping (sb200_pipeline_strip_dev).

Everything here is host logic over torch.distributed, so it runs unchanged on the gloo
backend with CPU tensors (tests/test_sharding_gloo.py) and on NCCL with CUDA tensors.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def batch_shard(n_pairs: int, rank: int, world: int) -> List[int]:
    """indices of the pairs rank `rank` processes: round-robin, no communication"""
    return list(range(rank, n_pairs, world))


def strip_bounds(height: int, world: int, align: int = 64) -> List[Tuple[int, int]]:
    """[y0, y1) of every rank's output strip.  Boundaries are multiples of `align` rows so that
    the same global rows start a strip no matter how many ranks there are."""
    units = (height + align - 1) // align
    base, extra = divmod(units, world)
    bounds, y = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * align
        y1 = min(height, y + n)
        bounds.append((y, y1))
        y = y1
    return bounds


def strip_geometry(height: int, rank: int, world: int, halo: int, align: int = 64) -> dict:
    """sb200_strip for this rank: rows held = halo_top + rows + halo_bot"""
    y0, y1 = strip_bounds(height, world, align)[rank]
    return dict(y0=y0, rows=y1 - y0, halo_top=min(halo, y0), halo_bot=min(halo, height - y1), frame_h=height)


def exchange_halo_rows(own: torch.Tensor, geom: dict, rank: int, world: int, halo: int, group=None) -> torch.Tensor:
    """own: this rank's rows [y0, y1) of an image, shape (rows, ...).  Returns the held rows
    [y0-halo_top, y1+halo_bot): the neighbours' boundary rows are fetched with one grouped
    send/recv per neighbour (dist.batch_isend_irecv -> ncclSend/ncclRecv on NCCL)."""
    top, bot = geom["halo_top"], geom["halo_bot"]
    rows = own.shape[0]
    if rows < halo and world > 1 and (top or bot):
        raise ValueError(f"strip of {rows} rows is shorter than the {halo}-row halo its neighbours need")
    recv_top = own.new_empty((top,) + tuple(own.shape[1:])) if top else None
    recv_bot = own.new_empty((bot,) + tuple(own.shape[1:])) if bot else None
    ops = []
    if top:  # rank-1 owns the rows above: it sends its last `top` rows, we send it our first `halo`
        ops.append(dist.P2POp(dist.irecv, recv_top, rank - 1, group))
        ops.append(dist.P2POp(dist.isend, own[:halo].contiguous(), rank - 1, group))
    if bot:
        ops.append(dist.P2POp(dist.isend, own[rows - halo:].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, recv_bot, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts = [t for t in (recv_top, own, recv_bot) if t is not None]
    return torch.cat(parts, 0) if len(parts) > 1 else own


def gather_strips(local: torch.Tensor, rank: int, world: int, height: int, align: int = 64, group=None):
    """optional collection of output strips on every rank (ncclAllGather of padded strips)"""
    bounds = strip_bounds(height, world, align)
    max_rows = max(y1 - y0 for y0, y1 in bounds)
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: y1 - y0] for o, (y0, y1) in zip(out, bounds)], 0)
