"""GPU parity for GpuTsdfMap against the CPU RayMapperTsdf oracle: the {weight, distance} layer bit for bit.

calculateTsdf is order dependent (clamped weighted mean per visit); the reference's own GPU kernel races on it and is
tested only on single rays (GpuTsdfTests.cpp:19-84).  Here free-space voxels (every visit far beyond the truncation
distance) are updated by count and surface voxels are replayed in ray order, so whole maps match exactly."""
import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from ohm_b200.lidar import LidarBox
from parity import compare_maps, integrate_both, make_pair

pytestmark = pytest.mark.gpu


def random_rays(n, extent, seed, origin=(0.05, 0.05, 0.05)):
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * n, 3))
    rays[0::2] = np.asarray(origin)
    rays[1::2] = rng.uniform(-extent, extent, size=(n, 3))
    return rays


def check_visits(g, c):
    gs, cs = g.stats(), c.stats()
    assert gs["rays_accepted"] == cs["rays_accepted"] and gs["voxel_visits"] == cs["voxel_visits"]


def test_tsdf_basic_biased_rays(gpu):
    # GpuTsdfTests.cpp:19-84 Tsdf.Basic: 16 biased rays, truncation 10 => distance == computeDistance along each ray
    bx, by, bz = 1.0, 0.9, 0.8
    ends = [(bx, 0, 0), (-bx, 0, 0), (bx, by, 0), (-bx, by, 0), (0, by, 0), (bx, -by, 0), (-bx, 0, bz), (bx, by, bz),
            (-bx, by, bz), (bx, 0, bz), (bx, -by, bz), (-bx, 0, bz), (bx, by, -bz), (-bx, by, -bz), (bx, 0, -bz),
            (bx, -by, -bz)]
    for e in ends:
        g, c = make_pair(0.1, mode="tsdf", origin=(-0.05, -0.05, -0.05), tsdf_trunc=10.0)
        integrate_both(g, c, np.array([[0, 0, 0], e], dtype=np.float64))
        compare_maps(g, c)
        check_visits(g, c)
        g.close()


def test_tsdf_random_rays_batched(gpu):
    g, c = make_pair(0.1, mode="tsdf")
    rays = random_rays(8192, 5.0, 3)
    integrate_both(g, c, rays, batch=2048)
    integrate_both(g, c, rays[::-1].copy().reshape(-1, 3), batch=4096)   # rays back from the samples
    compare_maps(g, c)
    check_visits(g, c)


@pytest.mark.parametrize("dims", [(12, 10, 6), (5, 7, 3), (16, 24, 8)])
def test_tsdf_region_dimensions(gpu, dims):
    g, c = make_pair(0.1, mode="tsdf", region_dim=dims, origin=(0.03, -0.07, 0.011))
    rays = random_rays(6000, 3.0, dims[1])
    integrate_both(g, c, rays, batch=2048)
    integrate_both(g, c, rays[::-1].copy().reshape(-1, 3))
    compare_maps(g, c)
    check_visits(g, c)


@pytest.mark.parametrize("kw", [dict(tsdf_trunc=0.3, tsdf_max_weight=20.0), dict(tsdf_dropoff=0.05),
                                dict(tsdf_sparsity=2.5), dict(tsdf_sparsity=0.0, tsdf_trunc=0.25)])
def test_tsdf_options(gpu, kw):
    g, c = make_pair(0.1, mode="tsdf", **kw)
    rng = np.random.RandomState(11)
    n = 4096
    rays = np.empty((2 * n, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    pts = rng.uniform(-4, 4, size=(n, 3))
    pts[: n // 2, 2] = -0.8 + rng.normal(scale=0.02, size=n // 2)   # a surface seen by many rays
    rays[1::2] = pts
    for _ in range(3):                                             # weights accumulate to max_weight
        integrate_both(g, c, rays, batch=1500)
    compare_maps(g, c)
    check_visits(g, c)


def test_tsdf_clip_filter_uses_unclipped_sample_for_distance(gpu):
    g, c = make_pair(0.1, mode="tsdf", filter_kind=gm.FILTER_CLIP_RANGE, filter_range=3.0)
    integrate_both(g, c, random_rays(2048, 6.0, 5))
    compare_maps(g, c)
    check_visits(g, c)


def test_tsdf_lidar_sweep(gpu):
    """BASELINE config 4 in miniature: one sweep into a 0.05 m TSDF map."""
    g, c = make_pair(0.05, mode="tsdf", device_bytes=24 << 30)
    rays, _, _ = LidarBox(1).sweep()
    sel = np.arange(0, rays.shape[0] // 2, 4)                      # every 4th ray: 32 k rays, ~14 M visits
    sub = np.empty((2 * len(sel), 3))
    sub[0::2], sub[1::2] = rays[2 * sel], rays[2 * sel + 1]
    integrate_both(g, c, sub)
    integrate_both(g, c, sub)
    compare_maps(g, c)
    check_visits(g, c)
