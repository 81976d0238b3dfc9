"""GPU parity for GpuNdtMap (NdtMode::kOccupancy) against the CPU RayMapperNdt oracle.

Bars: region set, visit/sample counts, voxel-mean coordinate and count, and the packed covariance are bit-exact
(the covariance update uses only + - * / sqrt in fp64, evaluated in the CPU order with FMA contraction off, and every
voxel's samples are replayed in ray order).  Log-odds carry a stated tolerance: the NDT miss term goes through exp()
and log(), whose last-ulp rounding differs between glibc and CUDA, and misses on a voxel are summed in a different
order — |gpu - cpu| <= 1e-5 + 1e-5 |cpu| (the reference's own GPU-vs-CPU bar is 1e-4, GpuNdtTests.cpp:162).
"""
import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from ohm_b200.lidar import LidarBox
from parity import check_counts, compare_maps, integrate_both, make_pair

pytestmark = pytest.mark.gpu

OCC_TOL = {gm.LAYER_OCCUPANCY: (1e-5, 1e-5)}


def test_ndt_hit_many_samples_one_voxel(gpu):
    # GpuNdtTests.cpp:171-228 Ndt.Hit: thousands of gaussian samples into one 2 m voxel, kRfExcludeRay
    g, c = make_pair(2.0, mode="ndt")
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    A = rng.uniform(-0.15, 0.15, size=(3, 3))
    samples = np.array([1.0, 1.0, 1.0]) + rng.normal(size=(6000, 3)) @ A.T
    samples = samples[np.all((samples > 0.01) & (samples < 1.99), axis=1)]
    rays = np.zeros((2 * len(samples), 3))
    rays[1::2] = samples
    integrate_both(g, c, rays, ray_flags=gm.RF_EXCLUDE_RAY, batch=2048)
    compare_maps(g, c)          # hits only: everything bit-exact, occupancy included
    st = check_counts(g, c)
    assert st["sample_updates"] == len(samples) and st["voxel_visits"] == 0


@pytest.mark.parametrize("shape", ["planar", "spherical"])
def test_ndt_miss_rays_through_gaussian_voxel(gpu, shape):
    # GpuNdtTests.cpp:235-406: populate one voxel, then fire rays through it; compare after every ray
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    n = 3000
    if shape == "planar":
        samples = np.column_stack([rng.uniform(0.01, 1.99, n), rng.uniform(0.01, 1.99, n), np.full(n, 1.0)])
        sensor = np.array([1.0, 1.0, 5.0])
        test_rays = [([1, 1, 5], [1, 1, -5]), ([1, 1, -5], [1, 1, 5]), ([-5, 1, 0.25], [5, 1, 0.25]),
                     ([1, 5, 1.01], [1, -5, 1.01]), ([-5, 1, 2], [5, 1, 1]), ([-5, 1, 2], [5, 1, 0.5])]
        origin = (0.0, 0.0, 0.0)
    else:
        r = 0.3
        v = rng.uniform(-0.99, 0.99, size=(n, 3))
        samples = v / np.linalg.norm(v, axis=1)[:, None] * rng.uniform(r - 0.05, r + 0.05, n)[:, None]
        sensor = np.array([0.0, 0.0, 5.0])
        test_rays = [([0, 0, 5], [0, 0, -5]), ([0, 0, -5], [0, 0, 5]), ([r, r, 5], [r, r, -5]),
                     ([1.5 * r, 1.5 * r, -5], [2 * r, 2 * r, 5])]
        origin = (-1.0, -1.0, -1.0)
    g, c = make_pair(2.0, mode="ndt", origin=origin)
    rays = np.zeros((2 * n, 3))
    rays[0::2] = sensor
    rays[1::2] = samples
    integrate_both(g, c, rays, ray_flags=gm.RF_EXCLUDE_RAY)
    compare_maps(g, c)
    for a, b_ in test_rays:
        integrate_both(g, c, np.array([a, b_], dtype=np.float64))
        compare_maps(g, c, tol_layers=OCC_TOL)
    check_counts(g, c)


def test_ndt_heavy_run_hits_and_misses_interleaved(gpu):
    """One voxel takes hundreds of samples AND hundreds of pass-through rays in the same batch, interleaved in ray
    order (the warp-per-run replay): every miss must see the Gaussian of its moment."""
    g, c = make_pair(2.0, mode="ndt")
    rng = np.random.RandomState(42)
    n = 1200
    rays = np.empty((2 * n, 3))
    for i in range(n):
        sensor = np.array([1.0, 1.0, 7.0]) + rng.uniform(-0.5, 0.5, 3)
        if i % 3 == 2:
            # passes through voxel (0,0,0) [0,2)^3 and ends two voxels below it
            end = np.array([rng.uniform(0.2, 1.8), rng.uniform(0.2, 1.8), -3.0])
        else:
            end = np.array([rng.uniform(0.05, 1.95), rng.uniform(0.05, 1.95), 1.0 + 0.05 * rng.normal()])
        rays[2 * i], rays[2 * i + 1] = sensor, end
    integrate_both(g, c, rays)           # one batch: ~800 hits + ~400 recorded misses on one voxel
    compare_maps(g, c, tol_layers=OCC_TOL)
    integrate_both(g, c, rays[::-1].reshape(-1, 2, 3)[:, ::-1].reshape(-1, 3))  # again, ray order reversed
    compare_maps(g, c, tol_layers=OCC_TOL)
    check_counts(g, c)


def test_ndt_random_rays_batched(gpu):
    g, c = make_pair(0.25, mode="ndt")
    rng = np.random.RandomState(7)
    rays = np.empty((2 * 8192, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    # samples on a few planes so that voxels accumulate enough points to become Gaussians, then get traversed
    pts = rng.uniform(-8, 8, size=(8192, 3))
    pts[:4096, 2] = -1.0 + rng.normal(scale=0.02, size=4096)
    pts[4096:6144, 0] = 6.0 + rng.normal(scale=0.02, size=2048)
    rays[1::2] = pts
    integrate_both(g, c, rays, batch=1024)
    integrate_both(g, c, rays[::-1].copy().reshape(-1, 3)[::1], batch=4096)   # reversed: rays from the samples back
    compare_maps(g, c, tol_layers=OCC_TOL)
    check_counts(g, c)


@pytest.mark.parametrize("dims", [(12, 10, 6), (5, 7, 3), (16, 24, 8)])
def test_ndt_region_dimensions(gpu, dims):
    # the Gaussian bits follow the counter tile's layout: byte copies when dim x is a multiple of 8, bit by bit otherwise
    g, c = make_pair(0.25, mode="ndt", region_dim=dims, origin=(0.3, -0.7, 0.11))
    rng = np.random.RandomState(dims[0])
    rays = np.empty((2 * 8192, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    pts = rng.uniform(-6, 6, size=(8192, 3))
    pts[:4096, 2] = -1.0 + rng.normal(scale=0.02, size=4096)
    pts[4096:6144, 0] = 4.0 + rng.normal(scale=0.02, size=2048)
    rays[1::2] = pts
    integrate_both(g, c, rays, batch=2048)
    integrate_both(g, c, rays[::-1].copy().reshape(-1, 3), batch=4096)
    compare_maps(g, c, tol_layers=OCC_TOL)
    check_counts(g, c)


def test_ndt_all_layers_with_timestamps(gpu):
    layers = [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_COVARIANCE, gm.LAYER_TOUCH_TIME, gm.LAYER_INCIDENT,
              gm.LAYER_TRAVERSAL]
    g, c = make_pair(0.2, mode="ndt", layers=layers)
    rng = np.random.RandomState(3)
    n = 6000
    rays = np.empty((2 * n, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    pts = rng.uniform(-5, 5, size=(n, 3))
    pts[: n // 2, 2] = -0.9 + rng.normal(scale=0.01, size=n // 2)
    rays[1::2] = pts
    ts = 5.0 + np.arange(n) * 2e-3
    integrate_both(g, c, rays, timestamps=ts, batch=2500)
    tol = dict(OCC_TOL)
    tol[gm.LAYER_TRAVERSAL] = (2e-5, 1e-6)
    compare_maps(g, c, tol_layers=tol)
    check_counts(g, c)


def test_ndt_lidar_two_sweeps(gpu):
    """BASELINE config 3 in miniature: moving-sensor sweeps at 0.1 m with voxel mean + covariance."""
    g, c = make_pair(0.1, mode="ndt", device_bytes=12 << 30)
    box = LidarBox(2)
    for _ in range(2):
        rays, _, ts = box.sweep()
        integrate_both(g, c, rays)
    compare_maps(g, c, tol_layers=OCC_TOL)
    st = check_counts(g, c)
    assert st["sample_updates"] == 2 * 131072


def test_ndt_tm_intensity_and_hit_miss(gpu):
    """NdtMode::kTraversability: intensity mean/covariance and hit/miss counts (CovarianceVoxelCompute.h:391-505)."""
    g, c = make_pair(0.25, mode="ndt_tm")
    assert g.params.ndt_tm == 1 and gm.LAYER_HIT_MISS in g.layers() and gm.LAYER_INTENSITY in g.layers()
    rng = np.random.RandomState(9)
    n = 4096
    rays = np.empty((2 * n, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    pts = rng.uniform(-6, 6, size=(n, 3))
    pts[:2048, 2] = -1.0 + rng.normal(scale=0.02, size=2048)
    rays[1::2] = pts
    intensities = rng.uniform(0, 255, n).astype(np.float32)
    for _ in range(3):
        integrate_both(g, c, rays, intensities=intensities, batch=1500)
    compare_maps(g, c, tol_layers=OCC_TOL)
    check_counts(g, c)
