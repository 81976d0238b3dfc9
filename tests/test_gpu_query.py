"""GPU parity for the read-side query on the resident map: ohm::RaysQuery / RaysQueryGpu (ohm/RaysQuery.cpp:109-199).

The query walks with the same fp64 DDA as the mappers, so against the CPU oracle (itself pinned bit-for-bit to the
reference's RaysQuery by tests/test_oracle_vs_ref.py) the bar is exact equality of ranges, unobserved volumes, terminal
states and terminal keys.  The reference's own GPU-vs-CPU test only asks for 1e-3 on ranges
(tests/ohmtestgpu/GpuRaysQueryTests.cpp).
"""
import numpy as np
import pytest

from ohm_b200 import gpumap as gm
from ohm_b200.lidar import LidarBox, cube_rays
from parity import integrate_both, make_pair

pytestmark = pytest.mark.gpu


def query_rays(n, seed, extent=9.0, origin_extent=2.0):
    rng = np.random.RandomState(seed)
    q = np.empty((2 * n, 3))
    q[0::2] = rng.uniform(-origin_extent, origin_extent, size=(n, 3))
    q[1::2] = rng.uniform(-extent, extent, size=(n, 3))
    q[4] = [np.nan, 0, 0]          # rejected by the map's ray filter
    q[9] = q[8]                    # degenerate ray: one voxel
    return q


def assert_same(a, b):
    for name, x, y in zip(("ranges", "unobserved volumes", "terminal states", "terminal keys"), a, b):
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8)), name


def test_rays_query_matches_cpu(gpu):
    g, c = make_pair(0.2)
    rng = np.random.RandomState(3)
    rays = np.empty((2 * 3000, 3))
    rays[0::2] = [0.05, 0.05, 0.05]
    rays[1::2] = rng.uniform(-7, 7, size=(3000, 3))
    integrate_both(g, c, np.concatenate([cube_rays(6000), rays]))
    q = query_rays(5000, 21)
    for coefficient in (1.0, 4.0 / 3.0 * np.pi * 1e-3):
        got, want = g.rays_query(q, coefficient), c.rays_query(q, coefficient)
        assert_same(got, want)
    states = got[2]
    assert (states == -2).sum() == 1 and (states == 1).sum() > 100 and (states == -1).sum() > 10


def test_rays_query_on_lidar_map_and_empty_map(gpu):
    g, c = make_pair(0.1)
    q = query_rays(2000, 5, extent=30.0)
    assert_same(g.rays_query(q), c.rays_query(q))          # empty map: everything unobserved, full-length ranges
    box = LidarBox(1)
    sweep, _, _ = box.sweep()
    integrate_both(g, c, sweep[: 2 * 40000])
    assert_same(g.rays_query(q), c.rays_query(q))
    # the query runs in stream order: queued behind a batch that is still in flight it sees that batch
    g.integrate_rays(sweep[2 * 40000: 2 * 60000])
    got = g.rays_query(q)
    c.integrate_rays(sweep[2 * 40000: 2 * 60000])
    assert_same(got, c.rays_query(q))


def test_line_keys_query_matches_cpu(gpu):
    """ohm::LineKeysQuery / LineKeysQueryGpu: keys along each line, both ends included, for an off-lattice origin and
    odd region dimensions (pure geometry: no map content involved)."""
    for kw in (dict(), dict(origin=(0.1, -0.2, 0.3), region_dim=(16, 24, 8))):
        g, c = make_pair(0.25, **kw)
        q = query_rays(1500, 33, extent=20.0)
        q[4] = [3.0, 0, 0]
        gi, gc, gk = g.line_keys_query(q)
        ci, cc, ck = c.line_keys_query(q)
        assert np.array_equal(gi, ci) and np.array_equal(gc, cc) and np.array_equal(gk, ck)
        assert gc.min() >= 1 and gk.shape[0] == int(gc.sum())
