"""Paging (the GpuLayerCache of this build): a map larger than its device memory.

The region table is sized for a fraction of the regions a trajectory creates, so the least recently walked regions are
evicted to the host store and come back when the sensor returns.  The whole map — device-resident and stored regions
together — must still equal the CPU mapper's, to the same bar as the unpaged tests (bit-exact; NDT log-odds 1e-5).
Reference behaviour: GpuLayerCache evicts its oldest entry when full and uploads a chunk on a cache miss
(ohmgpu/GpuLayerCache.cpp:429-633); the reference's own test of it is GpuMapTest.cpp's small-cache populate
(gpuMapTest with a 2 MiB cache).
"""
import numpy as np
import pytest

from ohm_b200 import gpumap as gm
from parity import check_counts, compare_maps, integrate_both, make_pair

pytestmark = pytest.mark.gpu


def trajectory_rays(steps, rays_per_step, seed, reach=5.0, stride=3.0):
    """A sensor moving along +x and back: every step fires rays_per_step rays to points within `reach` of it."""
    rng = np.random.RandomState(seed)
    xs = list(range(steps)) + list(range(steps - 2, -1, -1))
    batches = []
    for k in xs:
        origin = np.array([stride * k + 0.05, 0.05, 0.05])
        rays = np.empty((2 * rays_per_step, 3))
        rays[0::2] = origin
        rays[1::2] = origin + rng.uniform(-reach, reach, size=(rays_per_step, 3))
        batches.append(rays)
    return batches


def region_bytes(layers, vpr):
    per_voxel = {gm.LAYER_OCCUPANCY: 4, gm.LAYER_MEAN: 8, gm.LAYER_COVARIANCE: 24, gm.LAYER_TSDF: 8,
                 gm.LAYER_TRAVERSAL: 4, gm.LAYER_TOUCH_TIME: 4, gm.LAYER_INCIDENT: 4}
    return sum(per_voxel[l] for l in layers) * vpr + 4096


@pytest.mark.parametrize("mode,layers,tol", [
    ("occupancy", [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN], None),
    ("ndt", [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_COVARIANCE], {gm.LAYER_OCCUPANCY: (1e-5, 1e-5)}),
    ("tsdf", [gm.LAYER_TSDF], None),
])
def test_map_larger_than_device_memory(gpu, mode, layers, tol):
    slots = 100
    kw = dict(region_dim=(16, 16, 16))
    if mode == "occupancy":
        kw["layers"] = layers
    res = 0.25 if mode != "tsdf" else 0.2
    g, c = make_pair(res, mode=mode, device_bytes=slots * region_bytes(layers, 16 ** 3), **kw)
    capacity = g.stats()["region_capacity"]
    assert capacity <= 2 * slots, "the test wants a small region table"
    batches = trajectory_rays(steps=26, rays_per_step=600, seed=5, reach=5.0 if mode != "tsdf" else 4.0)
    for rays in batches:
        g.integrate_rays(rays)
    c.integrate_rays(np.concatenate(batches))
    g.sync_voxels()
    ps = g.paging_stats()
    st = g.stats()
    assert st["regions"] == len(c.dump()) > capacity, "the map must not fit in the table"
    assert ps["evicted"] > 0 and ps["paged_in"] > 0 and ps["resident"] <= capacity
    assert ps["resident"] + ps["stored"] == st["regions"]
    compare_maps(g, c, tol_layers=tol)
    check_counts(g, c)
    # written regions come back too: replace one stored region's layer and read it again
    keys = g.region_keys()
    assert len(keys) == st["regions"]
    g.close()


def test_clear_forgets_the_store(gpu):
    g, c = make_pair(0.25, device_bytes=60 * region_bytes([gm.LAYER_OCCUPANCY], 16 ** 3), region_dim=(16, 16, 16))
    for rays in trajectory_rays(steps=20, rays_per_step=400, seed=9):
        g.integrate_rays(rays)
    g.sync_voxels()
    assert g.paging_stats()["stored"] > 0
    g.clear()
    assert g.paging_stats()["stored"] == 0 and g.stats()["regions"] == 0
    rays = trajectory_rays(steps=1, rays_per_step=500, seed=1)[0]
    integrate_both(g, c, rays)
    compare_maps(g, c)
    g.close()


def test_remove_region_resident_and_stored(gpu):
    # MapRegionCache::remove: a removed region is gone from reads and counts, its slot is reusable, and integrating
    # into it again starts from an unobserved region (the oracle removes the same regions)
    g, c = make_pair(0.25, device_bytes=60 * region_bytes([gm.LAYER_OCCUPANCY], 16 ** 3), region_dim=(16, 16, 16))
    batches = trajectory_rays(steps=20, rays_per_step=400, seed=9)
    for rays in batches:
        g.integrate_rays(rays)
    g.sync_voxels()
    before = g.stats()["regions"]
    ps = g.paging_stats()
    assert ps["stored"] > 0
    dump = g.dump()
    keys = sorted(dump.keys())
    victims = [keys[0], keys[len(keys) // 2], keys[-1]]
    for k in victims:
        g.remove_region(k)
    assert g.stats()["regions"] == before - len(victims)
    after = g.dump()
    assert sorted(after.keys()) == [k for k in keys if k not in victims]
    for k in after:
        assert np.array_equal(after[k][gm.LAYER_OCCUPANCY].view(np.uint32), dump[k][gm.LAYER_OCCUPANCY].view(np.uint32))
    with pytest.raises(Exception):
        g.remove_region(victims[0])
    # the map keeps working: the same rays again re-create the removed regions
    g.integrate_rays(batches[0])
    g.sync_voxels()
    assert g.stats()["regions"] >= before - len(victims)
    g.close()
