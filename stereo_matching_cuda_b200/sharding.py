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
import ctypes
import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def batch_shard(n_pairs: int, rank: int, world: int) -> List[int]:
    """indices of the pairs rank `rank` processes: round-robin, no communication"""
    return list(range(rank, n_pairs, world))


def strip_rows(height: int, rank: int, world: int) -> Tuple[int, int]:
    """(y0, rows) of rank `rank`: the balanced split of sb200_strip_rows (include/stereo_b200.h) -- whole units of 4 rows
    (the fused kernel's row step), the remainder units dealt to the first ranks, the last strip takes the odd rows."""
    units = (height + 3) // 4
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    un = base + (1 if rank < extra else 0)
    y0 = min(height, 4 * u0)
    return y0, min(height, 4 * (u0 + un)) - y0


def strip_bounds(height: int, world: int, halo: int = 0) -> List[Tuple[int, int]]:
    """[y0, y1) of every rank's output strip.  With `halo` given, every rank raises the same error when the frame is
    too short for `world` strips of at least `halo` rows (a neighbour could not serve the halo)."""
    bounds = []
    for r in range(world):
        y0, rows = strip_rows(height, r, world)
        bounds.append((y0, y0 + rows))
    if world > 1 and min(y1 - y0 for y0, y1 in bounds) < max(halo, 1):
        raise ValueError(f"a {height}-row frame cannot be split into {world} strips of at least {max(halo, 1)} rows")
    return bounds


def strip_geometry(height: int, rank: int, world: int, halo: int) -> dict:
    """sb200_strip for this rank: rows held = halo_top + rows + halo_bot"""
    y0, y1 = strip_bounds(height, world, halo)[rank]
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


def gather_strips(local: torch.Tensor, rank: int, world: int, height: int, group=None):
    """optional collection of output strips on every rank (ncclAllGather of padded strips)"""
    bounds = strip_bounds(height, world)
    max_rows = max(y1 - y0 for y0, y1 in bounds)
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: y1 - y0] for o, (y0, y1) in zip(out, bounds)], 0)


class NcclComm:
    """A raw ncclComm_t for sb200_pipeline_strips_nccl, built with ctypes on the libnccl this process already has
    (torch's): rank 0 draws the unique id, torch.distributed broadcasts its 128 bytes, every rank calls
    ncclCommInitRank on its current CUDA device.  A C++ host passes its own communicator instead."""

    class _Id(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_ubyte * 128)]  # (c_char arrays would be cut at the first NUL byte)

    def __init__(self, rank: int, world: int, group=None, lib_path: str = None):
        self.lib = ctypes.CDLL(lib_path or os.environ.get("SB200_NCCL_LIB") or "libnccl.so.2")
        self.lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(self._Id)]
        self.lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, self._Id, ctypes.c_int]
        self.lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
        self.lib.ncclGetErrorString.restype = ctypes.c_char_p
        uid = self._Id()
        if rank == 0:
            self._ck(self.lib.ncclGetUniqueId(ctypes.byref(uid)))
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        t = torch.tensor(list(uid.internal), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0, group=group)
        raw = t.cpu().numpy().tobytes()
        ctypes.memmove(ctypes.byref(uid), raw, 128)
        self.handle = ctypes.c_void_p()
        self.rank, self.world = rank, world
        self._ck(self.lib.ncclCommInitRank(ctypes.byref(self.handle), world, uid, rank))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("NCCL: " + self.lib.ncclGetErrorString(rc).decode())

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.ncclCommDestroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
