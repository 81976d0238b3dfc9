"""The C restatement (oracle/ohm_oracle.c) against the REFERENCE ITSELF: oracle/_ref/libohm_ref.so is ohm's own
RayMapperOccupancy / RayMapperNdt / RayMapperTsdf + OccupancyMap, compiled unmodified from /root/reference by
`make -C oracle ref`.  Every layer of every region must be bit-identical on the same seeded rays.

Runs on CPU.  Skipped only where neither the library nor /root/reference exists.
"""
import numpy as np
import pytest

from ohm_b200.lidar import LidarBox, cube_rays
from oracle import pyoracle as po
from oracle import pyref as pr

pytestmark = pytest.mark.skipif(not pr.available(), reason="oracle/_ref not built and /root/reference absent")


def assert_identical(o, r):
    a, b = o.dump(), r.dump()
    assert sorted(a) == sorted(b)
    for key in a:
        assert sorted(a[key]) == sorted(b[key])
        for layer in a[key]:
            x, y = np.ascontiguousarray(a[key][layer]), np.ascontiguousarray(b[key][layer])
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), y.view(np.uint32)
            assert np.array_equal(x, y), f"layer {layer} of region {key}: {int((x != y).sum())} words differ"
    return len(a)


def random_rays(n, extent, seed):
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * n, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    rays[1::2] = rng.uniform(-extent, extent, size=(n, 3))
    return rays


def test_config1_occupancy():
    o, r = po.OracleMap(0.2), pr.ReferenceMap(0.2)
    rays = cube_rays(10000)
    assert o.integrate_rays(rays) == r.integrate_rays(rays) == 10000
    assert assert_identical(o, r) == 1


def test_all_sample_layers_with_timestamps():
    layers = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_TRAVERSAL, po.LAYER_TOUCH_TIME, po.LAYER_INCIDENT]
    o, r = po.OracleMap(0.2, layers=layers), pr.ReferenceMap(0.2, layers=layers)
    rays = random_rays(8192, 6.0, 5)
    ts = 100.0 + np.arange(8192) * 1e-3
    for s in range(0, 8192, 3000):
        e = min(8192, s + 3000)
        o.integrate_rays(rays[2 * s:2 * e], timestamps=ts[s:e])
        r.integrate_rays(rays[2 * s:2 * e], timestamps=ts[s:e])
    assert_identical(o, r)
    assert o.first_ray_time() == r.first_ray_time() == 100.0


# bit 1 = kRfStopOnFirstOccupied; 0x63 = ohm::ClearingPattern::kDefaultRayFlags (ohm/ClearingPattern.h:45)
@pytest.mark.parametrize("flags", [0, 1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5, 1 << 6, 1 << 7, (1 << 2) | (1 << 0),
                                   (1 << 1) | (1 << 0), 0x63])
def test_ray_flags(flags):
    o, r = po.OracleMap(0.25), pr.ReferenceMap(0.25)
    rays = random_rays(4096, 12.0, 7)
    for m in (o, r):
        m.integrate_rays(rays[:4096])
        m.integrate_rays(rays[4096:], ray_flags=flags)
        m.integrate_rays(rays[2048:6144], ray_flags=flags)
    assert_identical(o, r)


def test_filters_and_bad_rays():
    for kw in (dict(filter_kind=po.FILTER_CLIP_RANGE, filter_range=10.0), dict(filter_range=30.0),
               dict(filter_kind=po.FILTER_NONE)):
        o, r = po.OracleMap(0.25, **kw), pr.ReferenceMap(0.25, **kw)
        rays = random_rays(2048, 25.0, 11)
        if kw.get("filter_kind", po.FILTER_GOOD_RAY) != po.FILTER_NONE:
            rays[3] = [np.nan, 0, 0]
            rays[9] = [np.inf, 1, 1]
        rays[50] = rays[51] = [0.30001, 0.2, 0.1]
        o.integrate_rays(rays)
        r.integrate_rays(rays)
        assert_identical(o, r)


def test_region_dims_origin_and_parameters():
    kw = dict(region_dim=(16, 24, 20), origin=(0.3, -0.7, 0.11), saturate_min=1, saturate_max=1, min_value=-1.0,
              max_value=2.5, hit_value=0.9, miss_value=-0.35)
    o, r = po.OracleMap(0.2, **kw), pr.ReferenceMap(0.2, **kw)
    rays = random_rays(6000, 9.0, 13)
    for _ in range(3):
        o.integrate_rays(rays)
        r.integrate_rays(rays)
    assert_identical(o, r)


def test_line_walk_keys_and_ranges():
    o, r = po.OracleMap(0.1, origin=(0.05, 0.05, 0.05)), pr.ReferenceMap(0.1, origin=(0.05, 0.05, 0.05))
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    for i in range(600):
        s, e = rng.uniform(-4, 4, 3), rng.uniform(-4, 4, 3)
        if i % 3 == 0:
            e = np.round(e, 1)              # voxel-boundary end points
            s = np.round(s, 1)
        for flags in (0, 1, 2, 3):
            ko, no, xo = o.walk_segment(s, e, flags)
            kr, nr, xr = r.walk_segment(s, e, flags)
            assert np.array_equal(ko, kr) and np.array_equal(no.view(np.uint64), nr.view(np.uint64))
            assert np.array_equal(xo.view(np.uint64), xr.view(np.uint64))


def test_ndt_mean_covariance_and_log_odds():
    o = po.OracleMap(0.25, mode="ndt", layers=[po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_COVARIANCE])
    r = pr.ReferenceMap(0.25, mode="ndt")
    rng = np.random.RandomState(5)
    n = 8192
    rays = np.empty((2 * n, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    pts = rng.uniform(-8, 8, size=(n, 3))
    pts[:4096, 2] = -1.0 + rng.normal(scale=0.02, size=4096)
    rays[1::2] = pts
    for _ in range(3):                      # later passes run real NDT misses through the Gaussians
        o.integrate_rays(rays)
        r.integrate_rays(rays)
    assert_identical(o, r)


def test_ndt_tm_intensity_and_hit_miss():
    layers = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_COVARIANCE, po.LAYER_INTENSITY, po.LAYER_HIT_MISS]
    o = po.OracleMap(0.25, mode="ndt_tm", layers=layers, ndt_tm=1)
    r = pr.ReferenceMap(0.25, mode="ndt_tm", ndt_tm=1)
    rng = np.random.RandomState(9)
    n = 4096
    rays = np.empty((2 * n, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    pts = rng.uniform(-6, 6, size=(n, 3))
    pts[:2048, 2] = -1.0 + rng.normal(scale=0.02, size=2048)
    rays[1::2] = pts
    intensities = rng.uniform(0, 255, n).astype(np.float32)
    for _ in range(3):
        o.integrate_rays(rays, intensities=intensities)
        r.integrate_rays(rays, intensities=intensities)
    assert_identical(o, r)


def test_tsdf():
    kw = dict(tsdf_trunc=0.3, tsdf_max_weight=50.0)
    o, r = po.OracleMap(0.1, mode="tsdf", layers=[po.LAYER_TSDF], **kw), pr.ReferenceMap(0.1, mode="tsdf", **kw)
    rays = random_rays(4000, 5.0, 3)
    for _ in range(2):
        o.integrate_rays(rays)
        r.integrate_rays(rays)
    assert_identical(o, r)


def test_lidar_sweep_config2():
    o, r = po.OracleMap(0.1), pr.ReferenceMap(0.1)
    rays, _, _ = LidarBox(1).sweep()
    o.integrate_rays(rays)
    r.integrate_rays(rays)
    assert assert_identical(o, r) > 1000


def test_clip_box_filter():
    # tests/ohmtestgpu/GpuMapTest.cpp:640-647,762-771 ClipBox: clipBounded against an Aabb
    box = (-3.0, -2.0, -1.5, 4.0, 2.5, 1.0)
    kw = dict(filter_kind=po.FILTER_CLIP_BOX, clip_box=box)
    o, r = po.OracleMap(0.25, **kw), pr.ReferenceMap(0.25, **kw)
    rays = random_rays(4096, 9.0, 17)
    rng = np.random.RandomState(18)
    rays[0:2000:2] = rng.uniform(-9, 9, size=(1000, 3))   # also rays that start outside the box
    rays[10] = rays[11]                                   # degenerate
    o.integrate_rays(rays)
    r.integrate_rays(rays)
    assert_identical(o, r)


def query_rays(n, seed, extent=9.0):
    rng = np.random.RandomState(seed)
    q = np.empty((2 * n, 3))
    q[0::2] = rng.uniform(-2, 2, size=(n, 3))
    q[1::2] = rng.uniform(-extent, extent, size=(n, 3))
    q[4] = [np.nan, 0, 0]          # rejected by the filter: range 0, volume 0, kNull
    q[9] = q[8]                    # degenerate: one voxel
    return q


def test_rays_query():
    # ohm::RaysQuery (ohm/RaysQuery.cpp:109-199) on a populated map: range to the first occupied voxel, unobserved
    # volume, terminal state and key, for rays through free, occupied and unobserved space
    o, r = po.OracleMap(0.2), pr.ReferenceMap(0.2)
    rays = np.concatenate([cube_rays(6000), random_rays(2000, 7.0, 5)])
    o.integrate_rays(rays)
    r.integrate_rays(rays)
    q = query_rays(4000, 21)
    for coefficient in (1.0, 4.0 / 3.0 * np.pi * 1e-3):
        a, b = o.rays_query(q, coefficient), r.rays_query(q, coefficient)
        for name, x, y in zip(("ranges", "unobserved volumes", "terminal states", "terminal keys"), a, b):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), name
    states = a[2]
    assert (states == -2).sum() == 1 and (states == 1).sum() > 100 and (states == -1).sum() > 10 and (states == 0).sum() > 0


def test_line_keys_query():
    # ohm::LineKeysQuery (ohm/LineKeysQuery.cpp:103-123): the keys along each line, both end voxels included
    o, r = po.OracleMap(0.25, origin=(0.1, -0.2, 0.3)), pr.ReferenceMap(0.25, origin=(0.1, -0.2, 0.3))
    q = query_rays(600, 33, extent=20.0)
    q[4] = [3.0, 0, 0]  # no filter in this query: keep the inputs finite
    a, b = o.line_keys_query(q), r.line_keys_query(q)
    for name, x, y in zip(("result indices", "result counts", "keys"), a, b):
        assert np.array_equal(x, y), name
    assert a[1].min() >= 1 and a[1][4] == 1 and a[2].shape[0] == int(a[1].sum())


def test_secondary_samples():
    # ohm::RayMapperSecondarySample (ohm/RayMapperSecondarySample.cpp:37-74): Welford range statistics per voxel, in ray
    # order, beside the occupancy mapper on the same map; ranges beyond the u16 millimetre clamp included
    layers = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_SECONDARY]
    o, r = po.OracleMap(0.25, layers=layers), pr.ReferenceMap(0.25, layers=layers)
    rng = np.random.RandomState(2)
    rays = np.empty((2 * 6000, 3))
    rays[0::2] = rng.uniform(-1, 1, size=(6000, 3))
    rays[1::2] = rays[0::2] + rng.normal(scale=0.4, size=(6000, 3))
    rays[20:40:2] = rays[21:41:2] + [80.0, 0, 0]
    for chunk in (rays[:4000], rays[4000:]):
        assert o.integrate_secondary(chunk) == r.integrate_secondary(chunk) == chunk.shape[0] // 2
    o.integrate_rays(random_rays(500, 4.0, 3))
    r.integrate_rays(random_rays(500, 4.0, 3))
    assert_identical(o, r)
