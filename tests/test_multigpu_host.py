"""World-size-2 gloo test of bench.py's N > 1 host logic (runs on CPU, no GPU): the exchange handles are all-gathered
in rank order, the trajectory's sweeps are dealt to the ranks step by step, and the parity gate's union check accepts
a correctly sharded map and rejects a duplicated or misplaced region.  The device side of the exchange is covered by
tests/test_gpu_exchange.py (one process, world 1/2/4/8) and by bench.py --gpus N itself (its parity gate)."""
import os
import socket
import sys

import numpy as np
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

    import bench
    from ohm_b200 import _lib
    from ohm_b200.lidar import cube_rays
    from oracle import pyoracle as po

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    # 1. handles: 128 opaque bytes per rank, gathered in rank order (what ohmb200_exchange_connect expects)
    mine = bytes([rank + 1]) * C.sizeof(_lib.ExchangeHandle)
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    assert [h[0] for h in handles] == [r + 1 for r in range(world)] and all(len(h) == 128 for h in handles)

    # 2. the steps' shares, in rank order, are the trajectory
    steps = 5
    dealt = [bench.own_sweep_index(k, r, world) for k in range(steps) for r in range(world)]
    assert dealt == list(range(steps * world))

    # 3. the gate's union check.  Every rank integrates the whole (rank-ordered) batch with the CPU mapper and keeps the
    #    regions it owns — what its GPU map would hold after the exchange.
    shares = [cube_rays(400 + 90 * r, half_extent=9.0, origin=(0.05 + 0.5 * r, 0.05, 0.05), seed=11 + r) for r in range(world)]
    m = po.OracleMap(0.25, mode="ndt", layers=[0, 1, 5])
    for s in shares:
        m.integrate_rays(s)
    full = m.dump()
    lib = _lib.load()
    owner = lambda key: lib.ohmb200_region_owner((C.c_int16 * 3)(*key), world)
    dump = {key: layers for key, layers in full.items() if owner(key) == rank}
    assert len(dump) > 0
    dumps = [None] * world
    dist.gather_object(dump, dumps if rank == 0 else None, dst=0)
    if rank == 0:
        ok = bench.compare_union(dumps, full, world)
        assert ok["ok"] and ok["regions"] == len(full), ok
        # a region on two ranks, a region on the wrong rank, a region missing, a flipped bit: all rejected
        key = next(iter(dumps[1]))
        dup = [{**dumps[0], key: dumps[1][key]}, dumps[1]]
        assert not bench.compare_union(dup, full, world)["ok"]
        moved = [{**dumps[0], key: dumps[1][key]}, {k: v for k, v in dumps[1].items() if k != key}]
        assert not bench.compare_union(moved, full, world)["ok"]
        missing = [dumps[0], {k: v for k, v in dumps[1].items() if k != key}]
        assert not bench.compare_union(missing, full, world)["ok"]
        bad = {k: dict(v) for k, v in dumps[1].items()}
        mean = bad[key][1].copy()
        mean.flat[0] ^= 1
        bad[key][1] = mean
        assert not bench.compare_union([dumps[0], bad], full, world)["ok"]
    m.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_host_logic(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
