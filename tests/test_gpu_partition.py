"""`ohmb200_set_partition` at world 4 and 8 on one device: every "rank" is handed every ray and keeps the regions it owns
(the route a TSDF map takes across GPUs, and the fallback where peer memory is not available).  The union of the ranks'
maps must be the single-map (= CPU mapper) result for every mapper: occupancy with voxel mean and traversal (the sample's
traversal share needs the exit range of the ray's last walked voxel — computed in closed form by the rank that owns the
SAMPLE, whoever walked that voxel), NDT, and TSDF (far-visit counts, near bits and ordered replays per owner).
The routed exchange (work sharded too) is tests/test_gpu_exchange.py."""
import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from oracle import pyoracle as po
from parity import bits

pytestmark = pytest.mark.gpu


def random_rays(count, extent, seed):
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * count, 3))
    rays[0::2] = np.array([0.05, 0.05, 0.05]) + rng.uniform(-0.4, 0.4, size=(count, 3))
    rays[1::2] = rng.uniform(-extent, extent, size=(count, 3))
    return rays


def run(world, mode, resolution, batches, layers=None, tol=None):
    cls = {"occupancy": ohm_b200.GpuMap, "ndt": ohm_b200.GpuNdtMap, "tsdf": ohm_b200.GpuTsdfMap}[mode]
    kw = {} if layers is None else {"layers": layers}
    parts = []
    for r in range(world):
        p = cls(resolution, device_bytes=512 << 20, **kw)
        p.set_partition(r, world)
        parts.append(p)
    cpu = po.OracleMap(resolution, mode=mode, layers=int(parts[0].params.layers))
    for rays in batches:
        for p in parts:
            p.integrate_rays(rays)
        cpu.integrate_rays(rays)
    ref = cpu.dump()
    union = {}
    for r, p in enumerate(parts):
        p.sync_voxels()
        for key, data in p.dump().items():
            assert key not in union and p.region_owner(key, world) == r, key
            union[key] = data
    assert sorted(union) == sorted(ref)
    for key in ref:
        for layer in parts[0].layers():
            a, b = union[key][layer], ref[key][layer]
            if tol and layer in tol:
                rtol, atol = tol[layer]
                af = np.nan_to_num(a.astype(np.float64), posinf=1e30, neginf=-1e30)
                bf = np.nan_to_num(b.astype(np.float64), posinf=1e30, neginf=-1e30)
                assert (np.abs(af - bf) <= atol + rtol * np.abs(bf)).all(), (key, gm.LAYER_NAMES[layer])
            else:
                assert np.array_equal(bits(a), bits(b)), (key, gm.LAYER_NAMES[layer])
    cs = cpu.stats()
    assert sum(p.stats()["voxel_visits"] for p in parts) == cs["voxel_visits"]
    if mode != "tsdf":
        assert sum(p.stats()["sample_updates"] for p in parts) == cs["sample_updates"]
    for p in parts:
        p.close()
    return len(ref)


@pytest.mark.parametrize("world", [4, 8])
def test_partition_occupancy_mean_traversal(gpu, world):
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TRAVERSAL]
    batches = [random_rays(6000, 14.0, seed=1), random_rays(6000, 14.0, seed=2)]
    assert run(world, "occupancy", 0.25, batches, layers, tol={gm.LAYER_TRAVERSAL: (2e-5, 1e-6)}) > 50


@pytest.mark.parametrize("world", [4, 8])
def test_partition_ndt(gpu, world):
    rng = np.random.RandomState(4)
    batches = []
    for step in range(3):
        n = 5000
        rays = np.empty((2 * n, 3))
        rays[0::2] = np.array([0.3 * step, 0.1, 0.4]) + rng.uniform(-0.2, 0.2, size=(n, 3))
        rays[1::2] = np.stack([rng.uniform(3.0, 3.3, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n)], axis=1)
        far = rng.rand(n) < 0.3
        rays[1::2][far, 0] += rng.uniform(1.5, 4.0, far.sum())
        batches.append(rays)
    run(world, "ndt", 0.2, batches, tol={gm.LAYER_OCCUPANCY: (1e-5, 1e-5)})


@pytest.mark.parametrize("world", [4, 8])
def test_partition_tsdf(gpu, world):
    batches = [random_rays(5000, 6.0, seed=7), random_rays(5000, 6.0, seed=8)]
    run(world, "tsdf", 0.1, batches)
