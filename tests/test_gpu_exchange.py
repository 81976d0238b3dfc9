"""The routed multi-GPU exchange (include/ohmb200.h: ohmb200_exchange_*), `world` maps driven from one process on one
device: every rank filters and cuts only its own rays, segments and samples travel to the owner of their region, and the
union of the per-rank maps must equal the CPU mapper integrating rank 0's rays, then rank 1's, ... — to the same bar as
the single-GPU parity tests (bit-exact; traversal and NDT log-odds within their stated tolerances).

The peers here share a process, so their inboxes are reached by pointer; across processes the same arena is mapped
through its CUDA IPC handle (bench.py --gpus N, tests/test_multigpu_host.py for the host logic).
"""
import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from ohm_b200.lidar import LidarBox
from oracle import pyoracle as po
from parity import bits

pytestmark = pytest.mark.gpu

NDT_TOL = {gm.LAYER_OCCUPANCY: (1e-5, 1e-5), gm.LAYER_INTENSITY: (1e-5, 1e-5)}


def random_rays(count, extent, seed, origin=(0.05, 0.05, 0.05)):
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * count, 3))
    rays[0::2] = np.asarray(origin) + rng.uniform(-0.3, 0.3, size=(count, 3))
    rays[1::2] = rng.uniform(-extent, extent, size=(count, 3))
    return rays


def make_world(world, resolution, mode="occupancy", per=8192, device_bytes=512 << 20, **overrides):
    cls = {"occupancy": ohm_b200.GpuMap, "ndt": ohm_b200.GpuNdtMap, "ndt_tm": ohm_b200.GpuNdtMap}[mode]
    kw = dict(overrides)
    if mode == "ndt_tm":
        kw["traversability"] = True
    maps = [cls(resolution, device_bytes=device_bytes, **kw) for _ in range(world)]
    gm.open_exchange(maps, per)
    okw = dict(overrides)
    okw["layers"] = int(maps[0].params.layers)
    okw["ndt_tm"] = int(maps[0].params.ndt_tm)
    cpu = po.OracleMap(resolution, mode=mode, **okw)
    return maps, cpu


def check_union(maps, cpu, tol_layers=None):
    """No region on two ranks, every region on its owner, the union equal to the oracle; sums of the counters too."""
    world = len(maps)
    for m in maps:
        m.sync_voxels()
    dumps = [m.dump() for m in maps]
    ref = cpu.dump()
    union = {}
    for r, d in enumerate(dumps):
        for key, layers in d.items():
            assert key not in union, f"region {key} lives on two ranks"
            assert maps[0].region_owner(key, world) == r, f"region {key} is not on its owner"
            union[key] = layers
    assert sorted(union) == sorted(ref), (sorted(set(union) - set(ref))[:4], sorted(set(ref) - set(union))[:4])
    for key in ref:
        for layer in maps[0].layers():
            a, b = union[key][layer], ref[key][layer]
            if tol_layers and layer in tol_layers:
                rtol, atol = tol_layers[layer]
                af = np.nan_to_num(a.astype(np.float64), posinf=1e30, neginf=-1e30)
                bf = np.nan_to_num(b.astype(np.float64), posinf=1e30, neginf=-1e30)
                assert (np.abs(af - bf) <= atol + rtol * np.abs(bf)).all(), (key, gm.LAYER_NAMES[layer])
            else:
                assert np.array_equal(bits(a), bits(b)), (key, gm.LAYER_NAMES[layer])
    cs = cpu.stats()
    gs = [m.stats() for m in maps]
    for name in ("rays_accepted", "voxel_visits", "sample_updates"):
        assert sum(g[name] for g in gs) == cs[name], name
    return len(ref)


def split(rays, world, seed, extras=()):
    """Uneven shares of a batch, in order (rank r takes a contiguous piece; one rank may get nothing)."""
    n = rays.shape[0] // 2
    rng = np.random.RandomState(seed)
    cuts = np.sort(rng.randint(0, n + 1, size=world - 1)) if world > 1 else np.array([], dtype=int)
    bounds = [0] + [int(c) for c in cuts] + [n]
    out = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        out.append((rays[2 * lo:2 * hi],) + tuple(None if e is None else e[lo:hi] for e in extras))
    return out


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_exchange_occupancy_all_layers(gpu, world):
    """Occupancy + mean + traversal + touch time + incident normals over several steps (both inbox parities)."""
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TRAVERSAL, gm.LAYER_TOUCH_TIME, gm.LAYER_INCIDENT]
    maps, cpu = make_world(world, 0.25, layers=layers)
    t = 0
    for m in maps:
        m.set_first_ray_time(50.0)
    for step in range(3):
        n = 6000
        rays = random_rays(n, 14.0, seed=100 + step)
        ts = 50.0 + (t + np.arange(n)) * 1e-3
        t += n
        shares = split(rays, world, seed=step, extras=(None, ts))
        gm.exchange_step(maps, shares)
        for share in shares:       # the CPU mapper is called once per rank's batch, in rank order
            if share[0].shape[0]:
                cpu.integrate_rays(share[0], None, share[2])
    regions = check_union(maps, cpu, tol_layers={gm.LAYER_TRAVERSAL: (2e-5, 1e-6)})
    assert regions > 50


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["ndt", "ndt_tm"])
def test_exchange_ndt(gpu, world, mode):
    """NDT under the exchange: the rays are broadcast (every owner evaluates the Gaussian misses of every ray that
    crosses its regions); mean, count, covariance and hit/miss counts exact, log-odds / intensity within 1e-5."""
    maps, cpu = make_world(world, 0.2, mode=mode, layers=[gm.LAYER_TOUCH_TIME] if mode == "ndt_tm" else [])
    for m in maps:
        m.set_first_ray_time(0.0)
    rng = np.random.RandomState(5)
    t = 0
    for step in range(4):
        n = 5000
        # a wall patch hit again and again: Gaussians establish (>= 3 samples) and later rays pass through them
        rays = np.empty((2 * n, 3))
        rays[0::2] = np.array([0.3 * step, 0.1, 0.4]) + rng.uniform(-0.2, 0.2, size=(n, 3))
        rays[1::2] = np.stack([rng.uniform(3.0, 3.3, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n)], axis=1)
        far = rng.rand(n) < 0.3
        rays[1::2][far, 0] += rng.uniform(1.5, 4.0, far.sum())      # these pass through the wall's voxels
        intens = rng.uniform(0, 255, n).astype(np.float32)
        ts = (t + np.arange(n)) * 1e-3
        t += n
        shares = split(rays, world, seed=10 + step, extras=(intens, ts))
        gm.exchange_step(maps, shares)
        for share in shares:
            if share[0].shape[0]:
                cpu.integrate_rays(share[0], share[1], share[2])
    check_union(maps, cpu, tol_layers=NDT_TOL)


def test_exchange_lidar_sweeps_match_single_gpu_and_oracle(gpu):
    """BASELINE config 5 in miniature: 4 ranks, each bringing one sweep of the moving sensor per step, NDT."""
    world = 4
    per = 131072
    maps, cpu = make_world(world, 0.1, mode="ndt", per=per, device_bytes=3 << 30)
    box = LidarBox(2 * world)
    for step in range(2):
        shares = []
        for r in range(world):
            rays, _, _ = box.sweep()
            quarter = rays[:2 * 20000]            # a slice of each sweep keeps the CPU oracle to seconds
            shares.append((quarter, None, None))
            cpu.integrate_rays(quarter)
        gm.exchange_step(maps, shares)
    check_union(maps, cpu, tol_layers=NDT_TOL)


def test_exchange_flags_and_filters(gpu):
    """Ray flags and the clip filter travel with the step; kRfStopOnFirstOccupied is refused."""
    world = 4
    maps, cpu = make_world(world, 0.25, filter_kind=gm.FILTER_CLIP_RANGE, filter_range=9.0)
    rays = random_rays(8000, 12.0, seed=3)
    rays[5] = [np.nan, 0, 0]
    for flags in (0, gm.RF_EXCLUDE_ORIGIN | gm.RF_END_POINT_AS_FREE, gm.RF_EXCLUDE_OCCUPIED, gm.RF_EXCLUDE_SAMPLE):
        shares = split(rays, world, seed=flags, extras=(None, None))
        gm.exchange_step(maps, shares, ray_flags=flags)
        for share in shares:
            if share[0].shape[0]:
                cpu.integrate_rays(share[0], None, None, flags)
    check_union(maps, cpu)
    with pytest.raises(ohm_b200.OhmB200Error):
        maps[0].exchange_send(rays[:64], ray_flags=gm.RF_STOP_ON_FIRST_OCCUPIED)
    with pytest.raises(ohm_b200.OhmB200Error):
        maps[0].integrate_rays(rays[:64])          # plain integrate is refused while the exchange is open
    maps[0].exchange_close()
    maps[0].integrate_rays(rays[:64])              # and works again afterwards
    maps[0].sync_voxels()


def test_exchange_steps_replayed_as_graphs(gpu):
    """A step of the same shape as one met before is recorded and replayed as one CUDA graph (send + integrate); the
    step number lives on the device.  Ten identical steps, world = 1: recorded at steps 3 and 4 (one per inbox parity),
    replayed from then on — the map must still equal the CPU mapper's after every step count."""
    maps, cpu = make_world(1, 0.25, layers=[gm.LAYER_OCCUPANCY, gm.LAYER_MEAN], per=8192)
    rays = random_rays(8000, 10.0, seed=9)
    before = maps[0].stats()["kernel_launches"]
    for step in range(10):
        gm.exchange_step(maps, [(rays, None, None)])
        cpu.integrate_rays(rays)
    check_union(maps, cpu)
    st = maps[0].stats()
    assert st["batches"] == 10 and st["kernel_launches"] - before >= 10 * 10
    # parameters changed between steps: the recorded steps bake the old ones in and must be dropped, not replayed
    maps[0].set_params(hit_value=1.25, miss_value=-0.5)
    cpu.set_params(hit_value=1.25, miss_value=-0.5)
    for step in range(4):
        gm.exchange_step(maps, [(rays, None, None)])
        cpu.integrate_rays(rays)
    check_union(maps, cpu)
    # a different batch after the replays: back to plain launches, same map
    other = random_rays(5000, 10.0, seed=10)
    gm.exchange_step(maps, [(other, None, None)])
    cpu.integrate_rays(other)
    check_union(maps, cpu)
