"""Deterministic synthetic rectified pairs: integer-hash texture on a virtual
infinite plane, smoothed [1 2 1]/4 horizontally; right R[y][x]=tex(x,y), left L[y][x]=tex(x-delta(y),y)
with delta a 16-band staircase from 0 to D-1, so the left label is -delta.  Shared by tests and bench."""
import numpy as np


def _hash(x, y, c, seed):
    v = (x.astype(np.uint64) * np.uint64(73856093)) ^ (y.astype(np.uint64) * np.uint64(19349663)) ^ np.uint64(
        (c * 83492791) ^ (seed * 2654435761 & 0xFFFFFFFF))
    v &= np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    v = (v * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(13)
    v = (v * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    v ^= v >> np.uint64(16)
    return (v >> np.uint64(24)).astype(np.int64)


def _tex(x, y, c, seed):
    # coarse 4x4 blocks blended with fine noise, then [1 2 1]/4 along x on the virtual plane
    def raw(xx):
        coarse = _hash((xx + 100000) // 4, (y + 100000) // 4, c, seed)
        fine = _hash(xx + 100000, y + 100000, c + 7, seed)
        return (3 * coarse + fine) // 4
    return (raw(x - 1) + 2 * raw(x) + raw(x + 1)) // 4


def delta_rows(h, size_d):
    band = (np.arange(h) * 16) // h
    return (band * (size_d - 1)) // 15


def make_pair(w, h, size_d, channels=1, seed=0, y0=0, rows=None):
    """returns (left, right) uint8 arrays of shape (rows, w) or (rows, w, channels): rows
    [y0, y0+rows) of an h-row frame (default: the whole frame)"""
    rows = h - y0 if rows is None else rows
    y, x = np.meshgrid(np.arange(y0, y0 + rows, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
    dl = delta_rows(h, size_d)[y0:y0 + rows, None]
    chans_l, chans_r = [], []
    for c in range(channels):
        chans_r.append(_tex(x, y, c, seed))
        chans_l.append(_tex(x - dl, y, c, seed))
    L = np.stack(chans_l, -1).astype(np.uint8)
    R = np.stack(chans_r, -1).astype(np.uint8)
    if channels == 1:
        return np.ascontiguousarray(L[..., 0]), np.ascontiguousarray(R[..., 0])
    return np.ascontiguousarray(L), np.ascontiguousarray(R)
