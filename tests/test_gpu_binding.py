"""The reference-side binding, compiled and run: tests/binding/GpuMapB200.cpp implements ohm::GpuMap / GpuNdtMap /
GpuTsdfMap, ohm::GpuCache (an ohm::MapRegionCache) and gpumap::enableGpu/sync/gpuCache over the C ABI, behind the
reference's UNMODIFIED headers (ohmgpu/GpuMap.h:143-384, GpuNdtMap.h:63-110, GpuTsdfMap.h:37-80, ohm/MapRegionCache.h:28-72),
and links against ohm's own OccupancyMap (oracle/_ref/libohm_ref.so).  Here a real ohm::OccupancyMap is filled through
`ohm::GpuMap(&map)` -> integrateRays -> syncVoxels() and compared, MapChunk layer by MapChunk layer, with a second map
filled by ohm's CPU mapper — the test tests/ohmtestgpu/GpuMapTest.cpp:90-92 runs, with an exact bar.

The binding is built where /root/reference exists (tests/binding/Makefile, by __graft_entry__.build()) and travels to
the GPU box as oracle/_ref/libohm_b200_binding.so; without it these tests cannot run and say so.
"""
import ctypes as C
import os

import numpy as np
import pytest

from ohm_b200.lidar import LidarBox

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libohm_b200_binding.so")

# ohm::MapFlag (ohm/MapFlag.h)
VOXEL_MEAN, TRAVERSAL, TOUCH_TIME, INCIDENT_NORMAL = 1 << 0, 1 << 2, 1 << 3, 1 << 4


@pytest.fixture(scope="module")
def binding(gpu):
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libohm_b200_binding.so was not built (needs /root/reference at build time)")
    lib = C.CDLL(LIB)
    lib.binding_compare.restype = C.c_longlong
    lib.binding_compare.argtypes = [C.c_int, C.c_uint, C.c_double, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint,
                                    C.c_size_t, C.c_double, C.c_size_t, C.POINTER(C.c_ulonglong)]
    lib.binding_cache_and_api.restype = C.c_int
    lib.binding_cache_and_api.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
    lib.binding_last_message.restype = C.c_char_p
    return lib


def random_rays(count, extent, seed):
    rng = np.random.RandomState(seed)
    rays = np.empty((2 * count, 3))
    rays[0::2] = np.array([0.05, 0.05, 0.05]) + rng.uniform(-0.2, 0.2, size=(count, 3))
    rays[1::2] = rng.uniform(-extent, extent, size=(count, 3))
    return np.ascontiguousarray(rays)


def compare(lib, mode, flags, resolution, rays, intensities=None, timestamps=None, ray_flags=0, batch=0, tol=0.0):
    regions = C.c_ulonglong(0)
    ip = intensities.ctypes.data if intensities is not None else None
    tp = timestamps.ctypes.data if timestamps is not None else None
    bad = lib.binding_compare(mode, flags, resolution, rays.ctypes.data, rays.shape[0], ip, tp, ray_flags, batch, tol,
                              1 << 30, C.byref(regions))
    assert bad == 0, f"{bad} differing words: {lib.binding_last_message().decode()}"
    return regions.value


def test_gpumap_behind_ohm_headers_matches_raymapper_occupancy(binding):
    """ohm::GpuMap(&map): occupancy + voxel mean + traversal + touch time + incident normals, app-sized batches."""
    n = 20000
    rays = random_rays(n, 12.0, seed=1)
    ts = np.ascontiguousarray(10.0 + np.arange(n) * 1e-3)
    flags = VOXEL_MEAN | TOUCH_TIME | INCIDENT_NORMAL
    assert compare(binding, 0, flags, 0.25, rays, timestamps=ts, batch=4096) == 27
    assert compare(binding, 0, 0, 0.1, random_rays(6000, 5.0, seed=2)) >= 8


def test_gpumap_lidar_sweep(binding):
    rays, _, _ = LidarBox(1).sweep()
    rays = np.ascontiguousarray(rays[:2 * 40000])
    assert compare(binding, 0, VOXEL_MEAN, 0.1, rays) > 50


@pytest.mark.parametrize("mode", [1, 2])
def test_gpundtmap_matches_raymapper_ndt(binding, mode):
    """ohm::GpuNdtMap(&map, true, 2048, mem, NdtMode): mean / covariance (/ intensity, hit-miss) bit-exact, log-odds 1e-5."""
    rng = np.random.RandomState(3)
    n = 12000
    rays = np.empty((2 * n, 3))
    rays[0::2] = np.array([0.1, 0.1, 0.4]) + rng.uniform(-0.2, 0.2, size=(n, 3))
    rays[1::2] = np.stack([rng.uniform(3.0, 3.3, n), rng.uniform(-2.5, 2.5, n), rng.uniform(-1.5, 1.5, n)], axis=1)
    far = rng.rand(n) < 0.3
    rays[1::2][far, 0] += rng.uniform(1.5, 4.0, far.sum())
    rays = np.ascontiguousarray(rays)
    intens = np.ascontiguousarray(rng.uniform(0, 255, n).astype(np.float32))
    compare(binding, mode, 0, 0.2, rays, intensities=intens, batch=3000, tol=1e-5)


def test_gputsdfmap_matches_raymapper_tsdf(binding):
    compare(binding, 3, 0, 0.1, random_rays(8000, 6.0, seed=5), batch=2500)


def test_map_callbacks_reach_the_cache_and_the_rest_of_the_api(binding):
    """OccupancyMap::clear / cullRegionsOutside -> MapRegionCache::clear / remove on the device map; upload of chunks the
    map already held; setHitValue / setMissValue; syncVoxels(layers) and gpumap::sync(map, layer); an arbitrary
    std::function ray filter; findLayerCache / syncLayerTo."""
    rays = random_rays(12000, 10.0, seed=7)
    step = binding.binding_cache_and_api(rays.ctypes.data, rays.shape[0], 1 << 30)
    assert step == 0, f"step {step}: {binding.binding_last_message().decode()}"
