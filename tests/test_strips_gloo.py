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
