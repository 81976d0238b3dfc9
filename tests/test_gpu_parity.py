"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Mirrors tests/ohmtestgpu/GpuMapTest.cpp of the reference (gpuMapTest / compareMaps), but with a bit-exact bar
for keys, counts and occupancy instead of the reference's "99% of voxels within half a hit".
"""
import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from ohm_b200.lidar import LidarBox, cube_rays
from parity import check_counts, compare_maps, integrate_both, make_pair

pytestmark = pytest.mark.gpu


def random_rays(count, extent, seed=5489, origin=(0.05, 0.05, 0.05)):
    # GpuMapTest.cpp:341-351: sensor at (0.05,0.05,0.05), samples uniform in +-extent (mt19937 default seed)
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * count, 3))
    rays[0::2] = np.asarray(origin)
    rays[1::2] = rng.uniform(-extent, extent, size=(count, 3))
    return rays


def test_config1_single_region(gpu):
    """BASELINE config 1: 10k rays into a 0.2 m map, one region."""
    g, c = make_pair(0.2)
    integrate_both(g, c, cube_rays(10000))
    s = compare_maps(g, c)
    assert s["regions"] == 1
    st = check_counts(g, c)
    assert st["rays_accepted"] == 10000


def test_populate_tiny(gpu):
    # GpuMapTest.cpp PopulateTiny: two rays
    g, c = make_pair(0.25)
    rays = np.array([[0.3, 0, 0], [1.1, 0, 0], [-5.0, 0, 0], [0.3, 0, 0]])
    integrate_both(g, c, rays)
    compare_maps(g, c)
    check_counts(g, c)


def test_populate_small(gpu):
    # GpuMapTest.cpp:332-350 PopulateSmall: 64 rays +-50 m, batches of 32
    g, c = make_pair(0.25)
    integrate_both(g, c, random_rays(64, 50.0), batch=32)
    compare_maps(g, c)
    check_counts(g, c)


def test_populate_large_batched(gpu):
    # GpuMapTest.cpp:352-371 PopulateLarge (131072 rays +-25 m, batch 2048), at a quarter of the ray count
    g, c = make_pair(0.25, device_bytes=4 << 30)
    integrate_both(g, c, random_rays(32768, 25.0), batch=2048)
    compare_maps(g, c)
    st = check_counts(g, c)
    assert st["batches"] == 16


def test_compare_in_voxel_rays_then_clear(gpu):
    # GpuMapTest.cpp:525-630 GpuMap.Compare: degenerate rays inside every voxel of a 16^3 region => every voxel
    # == hit value exactly; then 16 clearing rays with a stronger miss value.
    g, c = make_pair(0.25, region_dim=(16, 16, 16))
    centres = []
    for z in range(16):
        for y in range(16):
            for x in range(16):
                centres.append(c.voxel_centre([0, 0, 0, x, y, z]))
    centres = np.asarray(centres)
    rays = np.repeat(centres, 2, axis=0)
    integrate_both(g, c, rays)
    compare_maps(g, c)
    occ = g.region_layer((0, 0, 0), gm.LAYER_OCCUPANCY)
    assert np.all(occ == np.float32(g.hit_value()))
    # miss value := -hit + miss (GpuMapTest.cpp:592-593), then clear the bottom slice along y
    new_miss = float(np.float32(-g.hit_value()) + np.float32(g.miss_value()))
    g.set_miss_value(new_miss)
    c.set_params(miss_value=new_miss)
    clear = []
    for x in range(16):
        clear.append(c.voxel_centre([0, 0, 0, x, 0, 0]))
        clear.append(c.voxel_centre([0, 0, 0, x, 15, 0]))
    integrate_both(g, c, np.asarray(clear))
    compare_maps(g, c)
    occ = g.region_layer((0, 0, 0), gm.LAYER_OCCUPANCY).reshape(16, 16, 16)  # [z, y, x]
    assert np.all(occ[0, :15, :] < 0) and np.all(occ[0, 15, :] > 0) and np.all(occ[1:] == np.float32(g.hit_value()))


@pytest.mark.parametrize("dims", [(12, 10, 6), (5, 7, 3), (16, 24, 8), (9, 32, 4)])
def test_region_dimensions_and_origin(gpu, dims):
    # counter tiles of every shape: odd rows (padded to an even stride), rows that are not a multiple of 8 (scalar
    # fold), slabs that are (the 128-bit fold); several work items per region (a hot region is split and folded by CAS)
    g, c = make_pair(0.25, region_dim=dims, origin=(0.3, -0.7, 0.11),
                     layers=[gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TRAVERSAL])
    rays = random_rays(12000, 9.0, seed=dims[0])
    integrate_both(g, c, rays[:16000])
    integrate_both(g, c, rays[16000:], batch=1500)
    compare_maps(g, c, tol_layers={gm.LAYER_TRAVERSAL: (2e-5, 1e-6)})
    check_counts(g, c)
    g2, c2 = make_pair(0.25, region_dim=dims, origin=(0.3, -0.7, 0.11))   # without traversal: the counting walker
    integrate_both(g2, c2, rays)
    compare_maps(g2, c2)
    check_counts(g2, c2)


@pytest.mark.parametrize("flags", [
    gm.RF_END_POINT_AS_FREE, gm.RF_EXCLUDE_ORIGIN, gm.RF_EXCLUDE_SAMPLE, gm.RF_EXCLUDE_RAY,
    gm.RF_EXCLUDE_UNOBSERVED, gm.RF_EXCLUDE_FREE, gm.RF_EXCLUDE_OCCUPIED, gm.RF_REVERSE_WALK,
    gm.RF_EXCLUDE_ORIGIN | gm.RF_END_POINT_AS_FREE,
    # kRfStopOnFirstOccupied: order-dependent across rays, integrated by the exact sequential kernel (integrateOrdered)
    gm.RF_STOP_ON_FIRST_OCCUPIED, gm.RF_STOP_ON_FIRST_OCCUPIED | gm.RF_END_POINT_AS_FREE,
    # ohm::ClearingPattern::kDefaultRayFlags (ohm/ClearingPattern.h:45)
    gm.RF_END_POINT_AS_FREE | gm.RF_STOP_ON_FIRST_OCCUPIED | gm.RF_EXCLUDE_FREE | gm.RF_EXCLUDE_UNOBSERVED,
])
def test_ray_flags(gpu, flags):
    g, c = make_pair(0.25)
    rays = random_rays(4096, 12.0, seed=7)
    # a first default pass so that the exclusion flags see free, occupied and unobserved voxels
    integrate_both(g, c, rays[:4096])
    integrate_both(g, c, rays[4096:], ray_flags=flags)
    integrate_both(g, c, rays[2048:6144], ray_flags=flags)
    compare_maps(g, c)
    check_counts(g, c)


def test_clip_range_filter(gpu):
    # clipRayFilter: long rays are clipped, flagged kRffClippedEnd and their end voxel takes a miss, not a hit
    g, c = make_pair(0.25, filter_kind=gm.FILTER_CLIP_RANGE, filter_range=10.0)
    integrate_both(g, c, random_rays(4096, 25.0, seed=11))
    compare_maps(g, c)
    st = check_counts(g, c)
    assert st["sample_updates"] < 4096


def test_bad_rays_are_dropped(gpu):
    # GpuMapTest.cpp:817-835 CheckBadRays + goodRayFilter: NaN/inf/too-long rays are skipped, tiny rays terminate
    g, c = make_pair(0.1, filter_range=100.0)
    rays = random_rays(256, 5.0, seed=3)
    rays[3] = [np.nan, 0, 0]
    rays[9] = [np.inf, 1, 1]
    rays[20] = [0, -np.inf, 1]
    rays[41] = [500.0, 0, 0]               # beyond filter_range
    rays[50] = rays[51] = [0.30001, 0.2, 0.1]  # zero length
    rays[61] = rays[60] + 1e-9             # sub-epsilon
    integrate_both(g, c, rays)
    compare_maps(g, c)
    st = check_counts(g, c)
    assert st["rays_accepted"] == 256 - 4


def test_all_sample_layers(gpu):
    """Voxel mean, incident normal, touch time (exact) and traversal (fp tolerance: summation order)."""
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TRAVERSAL, gm.LAYER_TOUCH_TIME, gm.LAYER_INCIDENT]
    g, c = make_pair(0.2, layers=layers)
    n = 8192
    rays = random_rays(n, 6.0, seed=21)
    ts = 100.0 + np.arange(n) * 1e-3
    integrate_both(g, c, rays, timestamps=ts, batch=3000)
    # traversal is a running fp32 sum of per-visit path lengths: the GPU adds in a different order than the
    # sequential mapper, so the bar is fp32 summation noise (relative 2e-5, a few ulp of a sum of thousands of terms)
    compare_maps(g, c, tol_layers={gm.LAYER_TRAVERSAL: (2e-5, 1e-6)})
    assert g.first_ray_time() == c.first_ray_time() == 100.0
    check_counts(g, c)


def test_many_samples_one_voxel(gpu):
    # GpuVoxelMeanTests / Ndt.Hit shape: thousands of samples into a single 2 m voxel — mean count must be exact
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_INCIDENT]
    g, c = make_pair(2.0, layers=layers)
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    n = 5000
    rays = np.zeros((2 * n, 3))
    rays[0::2] = [1.0, 1.0, 5.0]
    rays[1::2] = rng.uniform(0.01, 1.99, size=(n, 3))
    integrate_both(g, c, rays, batch=1024)
    compare_maps(g, c)
    key = c.voxel_key([1.0, 1.0, 1.0])
    mean = g.region_layer(tuple(key[:3]), gm.LAYER_MEAN)
    idx = key[3] + 32 * key[4] + 32 * 32 * key[5]
    assert mean[idx, 1] == n


def test_lidar_sweep_config2(gpu):
    """BASELINE config 2: one 64x2048 sweep at 0.1 m, occupancy only — bit-exact map, exact V/S/N."""
    g, c = make_pair(0.1, device_bytes=4 << 30)
    rays, _, _ = LidarBox(1).sweep()
    integrate_both(g, c, rays)
    s = compare_maps(g, c)
    st = check_counts(g, c)
    assert st["rays_accepted"] == rays.shape[0] // 2
    assert s["regions"] > 1000


def test_lidar_two_sweeps_with_mean(gpu):
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN]
    g, c = make_pair(0.1, device_bytes=6 << 30, layers=layers)
    box = LidarBox(2)
    for k in range(2):
        rays, _, _ = box.sweep()
        # app-sized batches for the second sweep (ohmapp/OhmAppCpu.h:52)
        integrate_both(g, c, rays, batch=None if k == 0 else 4096 * 8)
    compare_maps(g, c)
    check_counts(g, c)


def test_empty_and_ragged(gpu):
    g, c = make_pair(0.25)
    assert g.L.ohmb200_integrate(g.h, None, 0, None, None, 0) == 0
    rays = random_rays(3, 4.0)
    # element_count odd: the trailing origin without a sample is ignored (element_count / 2 rays)
    g.integrate_rays(rays[:5])
    c.integrate_rays(rays[:4])
    g.sync_voxels()
    compare_maps(g, c)
    assert g.region_count() == len(c.region_keys())


def test_clear_and_reuse(gpu):
    g, c = make_pair(0.25)
    g.integrate_rays(random_rays(512, 10.0, seed=5))
    g.sync_voxels()
    assert g.region_count() > 0
    g.clear()
    assert g.region_count() == 0
    integrate_both(g, c, random_rays(512, 10.0, seed=6))
    compare_maps(g, c)


def test_write_region_roundtrip(gpu):
    # GpuLayerCache::upload equivalent: host chunk -> device -> host
    g, c = make_pair(0.25)
    chunk = np.random.RandomState(0).uniform(-2, 3, size=32 ** 3).astype(np.float32)
    g.write_region((3, -2, 1), gm.LAYER_OCCUPANCY, chunk)
    back = g.region_layer((3, -2, 1), gm.LAYER_OCCUPANCY)
    assert np.array_equal(back, chunk)
    assert g.region_count() == 1


@pytest.mark.parametrize("mode", ["occupancy", "ndt", "tsdf"])
def test_repeatable_bit_for_bit(gpu, mode):
    """The same rays into a fresh map, six times: every layer of every region must come out bit-identical (a data
    race between the persistent CTAs that share a region shows up here long before it shows up against the oracle)."""
    rays = np.concatenate([cube_rays(6000), random_rays(3000, 9.0, seed=3)])
    cls = {"occupancy": ohm_b200.GpuMap, "ndt": ohm_b200.GpuNdtMap, "tsdf": ohm_b200.GpuTsdfMap}[mode]
    first = None
    for rep in range(6):
        g = cls(0.2, device_bytes=1 << 30)
        g.integrate_rays(rays[:8000])
        g.integrate_rays(rays[8000:])
        d = g.dump()
        g.close()
        if first is None:
            first = d
            continue
        assert sorted(d) == sorted(first)
        for key in d:
            for layer, arr in d[key].items():
                a, b = np.ascontiguousarray(arr), np.ascontiguousarray(first[key][layer])
                if mode == "ndt" and layer == gm.LAYER_OCCUPANCY:
                    # NDT log-odds: the order of float atomics on Gaussian voxels is not fixed
                    assert np.allclose(np.nan_to_num(a, posinf=1e30), np.nan_to_num(b, posinf=1e30), rtol=1e-5, atol=1e-5)
                else:
                    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (rep, key, gm.LAYER_NAMES[layer])


def test_async_download_is_a_snapshot(gpu):
    """ohmb200_read_regions_async: the chunks are those of the moment of the call (stream order), whatever is
    integrated while the copy runs; two downloads may be in flight; a key that is not resident is reported by the wait."""
    import torch

    g, c = make_pair(0.25)
    first, second = random_rays(3000, 12.0, seed=11), random_rays(3000, 12.0, seed=12)
    g.integrate_rays(first)
    g.sync_voxels()
    keys = g.region_keys()
    expect = g.region_layers(keys, gm.LAYER_OCCUPANCY)
    chunk = g.L.ohmb200_region_layer_bytes(g.h, gm.LAYER_OCCUPANCY)
    a = torch.empty(len(keys) * chunk, dtype=torch.uint8).pin_memory()
    b = torch.empty(len(keys) * chunk, dtype=torch.uint8).pin_memory()
    g.region_layers_async(keys, gm.LAYER_OCCUPANCY, a.data_ptr(), a.numel())
    g.integrate_rays(second)  # runs beside the copy, must not show in it
    g.region_layers_async(keys, gm.LAYER_OCCUPANCY, b.data_ptr(), b.numel())
    g.download_wait()
    got_a = a.numpy().view(np.float32).reshape(len(keys), -1)
    got_b = b.numpy().view(np.float32).reshape(len(keys), -1)
    assert np.array_equal(got_a.view(np.uint32), expect.view(np.uint32))
    after = g.region_layers(keys, gm.LAYER_OCCUPANCY)
    assert np.array_equal(got_b.view(np.uint32), after.view(np.uint32))
    assert not np.array_equal(got_a.view(np.uint32), got_b.view(np.uint32))
    c.integrate_rays(first)
    c.integrate_rays(second)
    compare_maps(g, c)
    missing = np.array([[99, 99, 99]], dtype=np.int16)
    g.region_layers_async(missing, gm.LAYER_OCCUPANCY, a.data_ptr(), a.numel())
    with pytest.raises(ohm_b200.OhmB200Error):
        g.download_wait()
    g.download_wait()  # the flag is cleared by the failing wait


def test_region_partition_union_matches_single_map(gpu):
    """Multi-GPU sharding on one device: two maps owning complementary region sets, fed the same rays, hold
    between them exactly the single-map (= oracle) result — no region on both, every visit applied once."""
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN]
    rays = random_rays(8192, 20.0, seed=13)
    g, c = make_pair(0.25, layers=layers)
    c.integrate_rays(rays)
    world = 2
    parts = []
    for r in range(world):
        p = ohm_b200.GpuMap(0.25, device_bytes=1 << 30, layers=layers)
        p.set_partition(r, world)
        p.integrate_rays(rays)
        p.sync_voxels()
        parts.append(p)
    dumps = [p.dump() for p in parts]
    assert not (set(dumps[0]) & set(dumps[1]))
    union = dict(dumps[0])
    union.update(dumps[1])
    ref = c.dump()
    assert sorted(union) == sorted(ref)
    for key in ref:
        assert g.region_owner(key, world) == (0 if key in dumps[0] else 1)
        for layer in layers:
            a, b = np.ascontiguousarray(union[key][layer]), np.ascontiguousarray(ref[key][layer])
            if a.dtype == np.float32:
                a, b = a.view(np.uint32), b.view(np.uint32)
            assert np.array_equal(a, b), (key, layer)
    cs = c.stats()
    assert sum(p.stats()["voxel_visits"] for p in parts) == cs["voxel_visits"]
    assert sum(p.stats()["sample_updates"] for p in parts) == cs["sample_updates"]


def test_clip_box_filter(gpu):
    # GpuMapTest.cpp ClipBox / ClipBoxCompare: clipBounded(Aabb) — rays cut at the box, clipped ends take a miss
    box = (-3.0, -2.0, -1.5, 4.0, 2.5, 1.0)
    g, c = make_pair(0.25, filter_kind=gm.FILTER_CLIP_BOX, clip_box=box)
    rays = random_rays(4096, 9.0, 17)
    rng = np.random.RandomState(18)
    rays[0:2000:2] = rng.uniform(-9, 9, size=(1000, 3))
    rays[10] = rays[11]
    integrate_both(g, c, rays, batch=1000)
    compare_maps(g, c)
    check_counts(g, c)


def test_secondary_sample_mapper(gpu):
    """ohm::RayMapperSecondarySample on the device map: Welford statistics of |secondary - primary| per voxel, samples of
    a voxel applied in ray order (many rays per voxel, several batches), beside the occupancy mapper."""
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_SECONDARY]
    g, c = make_pair(0.25, layers=layers)
    rng = np.random.RandomState(2)
    rays = np.empty((2 * 20000, 3))
    rays[0::2] = rng.uniform(-1.5, 1.5, size=(20000, 3))
    rays[1::2] = rays[0::2] + rng.normal(scale=0.4, size=(20000, 3))
    rays[20:40:2] = rays[21:41:2] + [80.0, 0, 0]      # beyond the u16 millimetre clamp
    for lo, hi in ((0, 14000), (14000, 14002), (14002, 40000)):
        assert g.integrate_secondary(rays[lo:hi]) == 2 * c.integrate_secondary(rays[lo:hi]) == hi - lo
    integrate_both(g, c, random_rays(2000, 4.0, seed=3))
    compare_maps(g, c)
    assert g.stats()["sample_updates"] == c.stats()["sample_updates"]
    plain, _ = make_pair(0.25)
    with pytest.raises(ohm_b200.OhmB200Error):
        plain.integrate_secondary(rays[:10])


def test_stop_on_first_occupied_all_layers(gpu):
    """kRfStopOnFirstOccupied with every sample layer: a stopped ray loses its sample — mean, touch time, incident normal
    and the sample's traversal share included (ohm/RayMapperOccupancy.cpp:183,234); traversal still accumulates along
    the stopped part of the walk.  The ordered kernel keeps the reference's per-call last_exit_range, so traversal is
    bit-exact here."""
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TRAVERSAL, gm.LAYER_TOUCH_TIME, gm.LAYER_INCIDENT]
    g, c = make_pair(0.25, layers=layers)
    n = 3000
    rays = random_rays(n, 8.0, seed=31)
    ts = 5.0 + np.arange(n) * 1e-3
    integrate_both(g, c, rays[:2 * 1500], timestamps=ts[:1500])
    for lo, hi in ((1500, 2400), (2400, 3000), (0, 1500)):
        g.integrate_rays(rays[2 * lo:2 * hi], timestamps=ts[lo:hi], ray_flags=gm.RF_STOP_ON_FIRST_OCCUPIED)
        c.integrate_rays(rays[2 * lo:2 * hi], timestamps=ts[lo:hi], ray_flags=gm.RF_STOP_ON_FIRST_OCCUPIED)
    g.sync_voxels()
    compare_maps(g, c, tol_layers={gm.LAYER_TRAVERSAL: (2e-5, 1e-6)})
    st = check_counts(g, c)
    assert st["sample_updates"] < 1500 + 3000      # some rays did stop
    # sharded maps cannot honour the flag (a ray's stop depends on regions of other owners): refused, not ignored
    p = ohm_b200.GpuMap(0.25, device_bytes=1 << 28)
    p.set_partition(0, 2)
    with pytest.raises(ohm_b200.OhmB200Error):
        p.integrate_rays(rays[:64], ray_flags=gm.RF_STOP_ON_FIRST_OCCUPIED)


def test_traversal_of_rays_that_walk_no_voxel(gpu):
    """A ray whose sensor and sample share a voxel walks nothing; the reference then subtracts the exit range the
    PREVIOUS ray of the call left behind from the sample's traversal share (last_exit_range is a variable of the
    whole integrateRays call, ohm/RayMapperOccupancy.cpp:79,190,309).  Reproduced: closed-form exit ranges in prepRays
    + one carry-forward scan."""
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_TRAVERSAL]
    g, c = make_pair(0.25, layers=layers)
    rng = np.random.RandomState(77)
    n = 4000
    rays = random_rays(n, 6.0, seed=41)
    inside = rng.choice(n, size=1200, replace=False)          # these rays stay inside one voxel
    centres = (np.floor(rng.uniform(-5, 5, size=(1200, 3)) / 0.25) + 0.5) * 0.25
    rays[2 * inside] = centres + rng.uniform(-0.1, 0.1, size=(1200, 3))
    rays[2 * inside + 1] = centres + rng.uniform(-0.1, 0.1, size=(1200, 3))
    rays[0] = rays[1] = centres[0]                            # the first ray of the call: nothing before it
    # one call on both sides: the carried value is per call
    g.integrate_rays(rays)
    c.integrate_rays(rays)
    g.integrate_rays(rays[::-1].copy())
    c.integrate_rays(rays[::-1].copy())
    g.sync_voxels()
    compare_maps(g, c, tol_layers={gm.LAYER_TRAVERSAL: (2e-5, 1e-6)})
    check_counts(g, c)


def test_segment_list_overflow_drops_the_batch_and_recovers(gpu):
    """Tiny regions and long rays cut into more segments than the batch's list holds (96 per ray): the batch is
    dropped WHOLE — no out-of-bounds access, no partial update — ohmb200_sync reports OHMB200_E_OVERFLOW once, the list
    is enlarged, and the same batch integrated again gives the oracle's map."""
    dims = (2, 2, 2)
    g, c = make_pair(0.1, region_dim=dims, device_bytes=256 << 20)
    rays = random_rays(7000, 40.0, seed=51)   # ~300 region crossings per ray: 2.1 M segments > 16384 x 96
    g.integrate_rays(rays)
    with pytest.raises(ohm_b200.OhmB200Error, match="segment"):
        g.sync_voxels()
    g.sync_voxels()                       # reported once
    occ = g.dump()
    assert all(np.all(np.isinf(v[gm.LAYER_OCCUPANCY])) for v in occ.values())   # nothing was applied
    g.clear()
    for _ in range(4):                    # the list doubles per overflow: 96 -> 192 -> 384 segments per ray
        g.integrate_rays(rays)
        try:
            g.sync_voxels()
            break
        except ohm_b200.OhmB200Error:
            g.clear()
    c.integrate_rays(rays)
    compare_maps(g, c)
