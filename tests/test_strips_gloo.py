"""CPU, world_size 2 over gloo: the N>1 path of bench.py / engine.strips.  Each rank builds the global
plan, takes its row strip, and the ranks together must own every output row exactly once, read only
rows inside their input strip, and partition the tile list -- no data-path collective is needed (the
all_gather below is only the test's own check)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, margin, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.grid import Georef, generate_patches, tile_plan
    from flair_for_aigle_b200.engine.strips import shard_rows
    from flair_for_aigle_b200.flair_zonal_detection.slicing import ownership_windows
    geo = Georef(700000.0, 6600000.0, 0.2, W, H)
    plan = tile_plan(generate_patches(512, margin, 0.2, geo), geo, 512, margin)
    own = ownership_windows(plan)
    sh = shard_rows(plan, own, 512, H, world)[rank]
    # rows this rank owns, painted locally
    owned = torch.zeros(H, dtype=torch.int32)
    cover = np.zeros((sh.out_r1 - sh.out_r0, W), np.int32)
    for o in sh.own:
        if o[1] > o[0] and o[3] > o[2]:
            cover[o[0]:o[1], o[2]:o[3]] += 1
    ok_local = bool((cover == 1).all())                       # its strip is tiled exactly once
    owned[sh.out_r0:sh.out_r1] = 1
    # every read stays inside the strip the rank holds (zero fill only outside the ZONE)
    rows_ok = bool(((sh.plan[:, 0] + 512 <= (sh.in_r1 - sh.in_r0)) | (sh.plan[:, 0] + 512 + sh.in_r0 > H)).all()) and \
        bool(((sh.plan[:, 0] >= 0) | (sh.plan[:, 0] + sh.in_r0 < 0)).all())
    gathered = [torch.zeros(H, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, owned)
    n_tiles = torch.tensor([len(sh.tile_idx)])
    dist.all_reduce(n_tiles)
    if rank == 0:
        total = torch.stack(gathered).sum(0)
        q.put((bool((total == 1).all()), int(n_tiles.item()) == plan.shape[0]))
    q.put((rank, ok_local, rows_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W,H,margin", [(1500, 4000, 64), (900, 2600, 40)])
def test_row_strips_partition_world2(W, H, margin):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, W, H, margin, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    glob = [r for r in results if len(r) == 2][0]
    assert glob == (True, True)
    for r in results:
        if len(r) == 3:
            assert r[1] and r[2], r


@pytest.mark.parametrize("W,H,margin,world", [(1500, 4000, 64, 2), (2100, 5200, 64, 4), (900, 2600, 40, 3), (1000, 8000, 128, 8)])
def test_shard_local_ownership_equals_global_ownership_on_owned_rows(W, H, margin, world):
    """bench.py's e2e leg hands inference_and_write() only the rank's rows of the GLOBAL tile table, on a raster that holds
    only the rank's input rows: the last-writer ownership is then recomputed inside the shard.  Claim: on the output rows the
    rank owns under the global plan, the shard-local owner of every pixel is the global owner (a global last writer that
    belongs to the shard is also the shard's last writer), so the rank's slice of the result is bit-identical to the 1-GPU
    raster.  Checked by painting tile indices with both ownership maps."""
    from oracle.grid import Georef, generate_patches, tile_plan
    from flair_for_aigle_b200.engine.strips import shard_rows
    from flair_for_aigle_b200.flair_zonal_detection.slicing import ownership_windows
    geo = Georef(700000.0, 6600000.0, 0.2, W, H)
    plan = tile_plan(generate_patches(512, margin, 0.2, geo), geo, 512, margin)
    own = ownership_windows(plan)
    owner_global = np.full((H, W), -1, np.int64)
    for i, o in enumerate(own):
        if o[1] > o[0] and o[3] > o[2]:
            assert (owner_global[o[0]:o[1], o[2]:o[3]] == -1).all()           # a partition
            owner_global[o[0]:o[1], o[2]:o[3]] = i
    assert (owner_global >= 0).all()
    covered = np.zeros(H, bool)
    for sh in shard_rows(plan, own, 512, H, world):
        if len(sh.tile_idx) == 0:
            continue
        # what the e2e leg does: the shard's tiles as their own plan on the strip raster (rows relative to in_r0)
        sub = plan[sh.tile_idx].copy()
        sub[:, 0] -= sh.in_r0
        sub[:, 2] -= sh.in_r0
        sub_own = ownership_windows(sub)
        owner_local = np.full((sh.in_r1 - sh.in_r0, W), -1, np.int64)
        for j, o in enumerate(sub_own):
            if o[1] > o[0] and o[3] > o[2]:
                owner_local[o[0]:o[1], o[2]:o[3]] = sh.tile_idx[j]
        a, b = sh.out_r0 - sh.in_r0, sh.out_r1 - sh.in_r0
        assert np.array_equal(owner_local[a:b], owner_global[sh.out_r0:sh.out_r1]), sh.rank
        assert not covered[sh.out_r0:sh.out_r1].any()
        covered[sh.out_r0:sh.out_r1] = True
        # every read of the shard's tiles stays inside the strip it holds (zero fill only outside the zone)
        assert (sub[:, 0] + 512 <= sh.in_r1 - sh.in_r0).all() or (plan[sh.tile_idx, 0] + 512 > H).any()
    assert covered.all()
