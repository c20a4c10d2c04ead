"""Row-strip sharding of a zone across GPUs: no data-path collective.

The tile grid is global (slicing.py enumerates x-outer / y-inner on the whole zone).  Tile ROWS
(the y index) are split into contiguous ranges, one per rank.  A rank reads the input rows its
tiles touch (its output rows plus a margin halo; neighbouring ranks read the same halo rows from
the source) and OWNS the output rows its tiles own under the last-writer rule, so the union of
the ranks' output strips is exactly the single-GPU raster -- bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class StripShard:
    rank: int
    tile_idx: np.ndarray      # indices (enumeration order) of this rank's tiles in the global plan
    in_r0: int                # global input rows [in_r0, in_r1) this rank must hold
    in_r1: int
    out_r0: int               # global output rows [out_r0, out_r1) this rank owns
    out_r1: int
    plan: np.ndarray          # local plan: row0 relative to in_r0, top_px relative to out_r0
    own: np.ndarray           # local ownership windows: rows relative to out_r0


def shard_rows(plan: np.ndarray, own: np.ndarray, patch: int, height: int, world: int) -> List[StripShard]:
    """Split the global plan into ``world`` row strips (contiguous ranges of distinct tile rows,
    balanced by tile count)."""
    n = plan.shape[0]
    tops = plan[:, 2]
    # distinct tile rows ordered north -> south (increasing top_px)
    row_keys = np.unique(tops)
    groups = np.array_split(np.arange(len(row_keys)), world)
    shards: List[StripShard] = []
    for rank, g in enumerate(groups):
        if len(g) == 0:
            shards.append(StripShard(rank, np.zeros(0, np.int64), 0, 0, 0, 0, np.zeros((0, 6), np.int32),
                                     np.zeros((0, 4), np.int32)))
            continue
        keys = row_keys[g]
        idx = np.flatnonzero(np.isin(tops, keys))          # keeps enumeration order
        p, o = plan[idx].copy(), own[idx].copy()
        in_r0 = int(max(p[:, 0].min(), 0))
        in_r1 = int(min(p[:, 0].max() + patch, height))
        live = (o[:, 1] > o[:, 0]) & (o[:, 3] > o[:, 2])
        out_r0 = int(o[live, 0].min()) if live.any() else 0
        out_r1 = int(o[live, 1].max()) if live.any() else 0
        p[:, 0] -= in_r0
        p[:, 2] -= out_r0
        o[:, 0] -= out_r0
        o[:, 1] -= out_r0
        shards.append(StripShard(rank, idx, in_r0, in_r1, out_r0, out_r1, p, o))
    return shards
