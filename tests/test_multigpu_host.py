"""World-size-2 gloo test of the N>1 host logic (runs on CPU): ray slices are padded and all-gathered into the
full sweep every rank integrates, and the region-ownership function splits the oracle's regions into disjoint
per-rank sets whose union is the whole map."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import ctypes as C
    from ohm_b200 import _lib
    from ohm_b200.lidar import cube_rays
    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rays = cube_rays(1001, half_extent=20.0)          # ragged: 1001 rays over 2 ranks
    n = rays.shape[0] // 2
    per = (n + world - 1) // world
    pad = per * world
    padded = np.full((2 * pad, 3), np.nan)
    padded[:2 * n] = rays
    mine = torch.from_numpy(padded[2 * per * rank:2 * per * (rank + 1)].copy())
    full = torch.empty((2 * pad, 3), dtype=torch.float64)
    dist.all_gather_into_tensor(full, mine)
    full = full.numpy()
    assert np.array_equal(full[:2 * n], rays) and np.all(np.isnan(full[2 * n:]))

    # The oracle drops the NaN padding through the good-ray filter, exactly as the device filter does.
    m = po.OracleMap(0.25)
    m.integrate_rays(full)
    assert m.stats()["rays_accepted"] == n
    lib = _lib.load()
    keys = m.region_keys()
    owned = [tuple(int(v) for v in k) for k in keys
             if lib.ohmb200_region_owner(np.ascontiguousarray(k).ctypes.data_as(C.POINTER(C.c_int16)), world) == rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, owned)
    if rank == 0:
        union = [k for part in gathered for k in part]
        assert len(union) == len(set(union)) == len(keys)
        assert all(len(part) > 0 for part in gathered)
    dist.barrier()
    dist.destroy_process_group()


def _weak_worker(rank, world, port, out_dir):
    """bench.py's N>1 arm: rank r brings its own batch (sweep r; ragged sizes), padded with NaN rays to the longest;
    the all-gather in rank order is the batch every rank integrates."""
    sys.path.insert(0, ROOT)
    from ohm_b200.lidar import cube_rays
    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sweeps = [cube_rays(700 + 150 * r, half_extent=12.0, origin=(0.05 + 0.5 * r, 0.05, 0.05), seed=11 + r)
              for r in range(world)]
    per = max(s.shape[0] // 2 for s in sweeps)
    mine = np.full((2 * per, 3), np.nan)
    mine[:sweeps[rank].shape[0]] = sweeps[rank]
    full = torch.empty((2 * per * world, 3), dtype=torch.float64)
    dist.all_gather_into_tensor(full, torch.from_numpy(mine))
    full = full.numpy()
    for r in range(world):
        got = full[2 * per * r:2 * per * (r + 1)]
        n_r = sweeps[r].shape[0]
        assert np.array_equal(got[:n_r], sweeps[r]) and np.all(np.isnan(got[n_r:]))
    # the padded, gathered batch and the plain concatenation of the sweeps are the same batch to the mapper
    a, b = po.OracleMap(0.25), po.OracleMap(0.25)
    a.integrate_rays(full)
    b.integrate_rays(np.concatenate(sweeps))
    assert a.stats()["rays_accepted"] == b.stats()["rays_accepted"] == sum(s.shape[0] // 2 for s in sweeps)
    da, db = a.dump(), b.dump()
    assert sorted(da) == sorted(db)
    for key in da:
        for layer in da[key]:
            assert np.array_equal(np.ascontiguousarray(da[key][layer]).view(np.uint8),
                                  np.ascontiguousarray(db[key][layer]).view(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_one_sweep_per_rank(tmp_path):
    world = 2
    mp.spawn(_weak_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)


def test_two_rank_gloo_gather_and_partition(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
