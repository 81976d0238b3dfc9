"""Pins the CPU oracle against the known-answer material the reference's own tests hold for the hot path
(SURVEY.md §8c).  The reference stores no golden outputs for ray integration; its tests are exact-value checks,
invariants and independent re-derivations.  Each test below restates one of them against oracle/ohm_oracle.c.

Runs on CPU (no GPU marker).
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import pyoracle as po

HIT = np.float32(math.log(0.9 / 0.1))
INF = np.float32(np.inf)


def f32(x):
    return C.c_float(float(x))


# -- tests/ohmtest/MapTests.cpp:34-79 --------------------------------------------------------------------------
def test_map_hit_and_miss_exact_values():
    L = po.lib()
    p = po.default_params(0.25)
    v = C.c_float(0)
    L.oracle_occupancy_adjust_hit(C.byref(v), INF, p.hit_value, INF, p.max_value, f32(-3.4e38), f32(3.4e38), 0)
    assert np.float32(v.value) == np.float32(p.hit_value) and v.value > 0          # EXPECT_EQ(value, map.hitValue())
    L.oracle_occupancy_adjust_miss(C.byref(v), INF, p.miss_value, INF, p.min_value, f32(-3.4e38), f32(3.4e38), 0)
    assert np.float32(v.value) == np.float32(p.miss_value) and v.value < 0         # EXPECT_EQ(value, map.missValue())
    # default probabilities (OccupancyMap.cpp:209-213): hit 0.9, miss 0.45, clamp [-2, 3.511]
    assert abs(p.hit_value - 2.1972246) < 1e-6 and abs(p.miss_value + 0.2006707) < 1e-6
    assert p.min_value == -2.0 and abs(p.max_value - 3.511) < 1e-6 and p.threshold_value == 0.0


def test_occupancy_clamps_and_saturation():
    L = po.lib()
    p = po.default_params(0.1)
    v = C.c_float(0)
    lo, hi = f32(-3.4e38), f32(3.4e38)
    # repeated hits saturate at max_value, repeated misses at min_value
    cur = INF
    for _ in range(5):
        L.oracle_occupancy_adjust_hit(C.byref(v), f32(cur), p.hit_value, INF, p.max_value, lo, hi, 0)
        cur = np.float32(v.value)
    assert cur == np.float32(p.max_value)
    for _ in range(40):
        L.oracle_occupancy_adjust_miss(C.byref(v), f32(cur), p.miss_value, INF, p.min_value, lo, hi, 0)
        cur = np.float32(v.value)
    assert cur == np.float32(p.min_value)
    # saturation: a voxel at the saturation bound is not modified (VoxelOccupancyCompute.h:25-31)
    L.oracle_occupancy_adjust_hit(C.byref(v), f32(p.min_value), p.hit_value, INF, p.max_value, f32(p.min_value), hi, 0)
    assert np.float32(v.value) == np.float32(p.min_value)
    # null update leaves an unobserved voxel unobserved
    L.oracle_occupancy_adjust_miss(C.byref(v), INF, p.miss_value, INF, p.min_value, lo, hi, 1)
    assert np.isinf(v.value)


# -- tests/ohmtestgpu/GpuMapTest.cpp:525-630 (GpuMap.Compare) --------------------------------------------------
def test_in_voxel_rays_give_exactly_hit_value_then_clear():
    m = po.OracleMap(0.25, region_dim=(16, 16, 16))
    centres = np.asarray([m.voxel_centre([0, 0, 0, x, y, z]) for z in range(16) for y in range(16) for x in range(16)])
    m.integrate_rays(np.repeat(centres, 2, axis=0))
    occ = m.region_layer((0, 0, 0), po.LAYER_OCCUPANCY)
    assert len(m.region_keys()) == 1 and np.all(occ == np.float32(m.params.hit_value))
    new_miss = float(np.float32(-m.params.hit_value) + np.float32(m.params.miss_value))
    m.set_params(miss_value=new_miss)
    clear = []
    for x in range(16):
        clear += [m.voxel_centre([0, 0, 0, x, 0, 0]), m.voxel_centre([0, 0, 0, x, 15, 0])]
    m.integrate_rays(np.asarray(clear))
    occ = m.region_layer((0, 0, 0), po.LAYER_OCCUPANCY).reshape(16, 16, 16)
    assert np.all(occ[0, :15, :] == np.float32(np.float32(m.params.hit_value) + np.float32(new_miss)))
    assert np.all(occ[0, 15, :] == np.float32(min(2 * np.float32(m.params.hit_value), np.float32(m.params.max_value))))
    assert np.all(occ[1:] == np.float32(m.params.hit_value))


# -- tests/ohmtest/KeyTests.cpp: key quantisation ---------------------------------------------------------------
def test_voxel_key_roundtrip_and_region_span():
    m = po.OracleMap(0.1)
    rng = np.random.RandomState(0)
    for p in rng.uniform(-20, 20, size=(2000, 3)):
        key = m.voxel_key(p)
        assert key is not None and np.all(key[3:] >= 0) and np.all(key[3:] < 32)
        c = m.voxel_centre(key)
        assert np.all(np.abs(c - p) <= 0.05 + 1e-9)
        assert np.array_equal(m.voxel_key(c), key)
    # region k spans [(k - 1/2) R, (k + 1/2) R): ohm/MapCoord.h:85-93
    assert list(m.voxel_key([1.59, 0, 0])[:1]) == [0] and list(m.voxel_key([1.61, 0, 0])[:1]) == [1]
    assert list(m.voxel_key([-1.61, 0, 0])[:1]) == [-1]
    assert list(m.voxel_key([0.05, 0.05, 0.05])) == [0, 0, 0, 16, 16, 16]


# -- tests/ohmtest/LineWalkTests.cpp:43-197 (testWalk) ---------------------------------------------------------
def _range_between(a, b, dim=32):
    return (b[3:] - a[3:]) + (b[:3] - a[:3]) * dim


def _ray_hits_box(start, end, lo, hi):
    d = end - start
    n = np.linalg.norm(d)
    if n == 0:
        return bool(np.all(start >= lo) and np.all(start <= hi))
    d = d / n
    tmin, tmax = -np.inf, np.inf
    for a in range(3):
        if d[a] == 0:
            if start[a] < lo[a] or start[a] > hi[a]:
                return False
            continue
        t0, t1 = (lo[a] - start[a]) / d[a], (hi[a] - start[a]) / d[a]
        tmin, tmax = max(tmin, min(t0, t1)), min(tmax, max(t0, t1))
    return tmax >= max(tmin, 0.0) - 1e-12


def _check_walk(m, start, end, include_end):
    skey, ekey = m.voxel_key(start), m.voxel_key(end)
    keys, enter, exit_ = m.walk_segment(start, end, 0 if include_end else 2)
    res = m.params.resolution
    last = None
    last_dist = -1.0
    for i, k in enumerate(keys):
        to_end = _range_between(k, ekey)
        dist = float(np.linalg.norm(to_end))
        if i == 0:
            assert np.array_equal(k, skey)                                   # first voxel is the start key
        else:
            assert not np.array_equal(k, skey)
            assert abs(np.linalg.norm(_range_between(last, k)) - 1.0) < 1e-6  # one orthogonal step
            assert dist < last_dist                                           # monotone approach
        last_dist = dist
        c = m.voxel_centre(k)
        pad = 0.5 * (res + 1e-3)
        assert _ray_hits_box(np.asarray(start), np.asarray(end), c - pad, c + pad)
        assert enter[i] <= exit_[i] + 1e-12
        last = k
    if include_end:
        assert np.array_equal(last, ekey)                                     # end voxel reported last
        assert len(keys) == 1 + int(np.abs(_range_between(skey, ekey)).sum())  # 1 + |dx|+|dy|+|dz|
        assert abs(exit_[-1] - np.linalg.norm(np.asarray(end) - np.asarray(start))) < 1e-12 or len(keys) == 1
    elif not np.array_equal(skey, ekey):
        assert abs(np.linalg.norm(_range_between(last, ekey)) - 1.0) < 1e-6
    else:
        assert len(keys) == 0


def test_line_walk_random():
    # LineWalk.Random: 1000 seeded rays in [-1,1]^3, two map origins, both end-point modes
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    for origin in ((0.0, 0.0, 0.0), (0.05, 0.05, 0.05)):
        m = po.OracleMap(0.1, origin=origin)
        for _ in range(1000):
            s, e = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 3)
            _check_walk(m, s, e, True)
            _check_walk(m, s, e, False)


def test_line_walk_lattice():
    # LineWalk.Walk: 27 directions x 10 scales from the origin — exact voxel-boundary walking
    for origin in ((0.0, 0.0, 0.0), (0.05, 0.05, 0.05)):
        m = po.OracleMap(0.1, origin=origin)
        for s in range(1, 11):
            for z in (-1, 0, 1):
                for y in (-1, 0, 1):
                    for x in (-1, 0, 1):
                        end = np.array([x, y, z], dtype=np.float64)
                        if s != 10:
                            end = end * (s / 10.0)
                        _check_walk(m, np.zeros(3), end, True)
                        _check_walk(m, np.zeros(3), end, False)


def test_degenerate_rays_terminate():
    # GpuMapTest.cpp:817-835 CheckBadRays: a sub-epsilon ray must not hang and reports only its own voxel
    m = po.OracleMap(0.1)
    keys, _, _ = m.walk_segment([0.31, 0.2, 0.1], [0.31 + 1e-9, 0.2, 0.1])
    assert len(keys) == 1
    keys, _, _ = m.walk_segment([0.31, 0.2, 0.1], [0.31, 0.2, 0.1], 2)
    assert len(keys) == 0


# -- tests/ohmtest/VoxelMeanTests.cpp (tolerance resolution / 1000) ---------------------------------------------
def test_voxel_mean_quantisation_and_progressive_mean():
    L = po.lib()
    res = 0.5
    out = np.zeros(3)
    dp = out.ctypes.data_as(C.POINTER(C.c_double))
    for p in (0.0, 0.05, 0.15, 0.20, 0.25, 0.30, 0.35, 0.40, 0.45, 0.50):
        local = np.full(3, p - 0.25)
        coord = L.oracle_sub_voxel_update(0, 0, local.ctypes.data_as(C.POINTER(C.c_double)), res)
        assert coord & (1 << 31)
        L.oracle_sub_voxel_to_local(coord, res, dp)
        assert np.all(np.abs(out - local) <= res / 1e3)
    rng = np.random.RandomState(3)
    pts = rng.uniform(-0.25, 0.25, size=(500, 3))
    coord, count = 0, 0
    for p in pts:
        coord = L.oracle_sub_voxel_update(coord, count, np.ascontiguousarray(p).ctypes.data_as(C.POINTER(C.c_double)), res)
        count += 1
    L.oracle_sub_voxel_to_local(coord, res, dp)
    # progressive quantised mean drifts by at most a few quantisation steps (res / 1023 each)
    assert np.all(np.abs(out - pts.mean(axis=0)) < 10 * res / 1023)


# -- tests/ohmtest/IncidentsTests.cpp + GpuIncidentsTests.cpp:111 (tol 1e-2) ------------------------------------
def test_incident_normal_running_mean():
    L = po.lib()
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    packed = 0
    dirs = []
    for r in range(1000):
        o = rng.uniform(-1, 1, 3)
        o = o / np.linalg.norm(o) * 0.3
        o[2] = abs(o[2])
        dirs.append(o / np.linalg.norm(o))
        inc = o.astype(np.float32)
        packed = L.oracle_update_incident_normal(packed, inc.ctypes.data_as(C.POINTER(C.c_float)), r)
    n = np.zeros(3, dtype=np.float32)
    L.oracle_decode_normal(packed, n.ctypes.data_as(C.POINTER(C.c_float)))
    assert abs(np.linalg.norm(n) - 1.0) < 1e-2
    # decode(encode(v)) ~= v
    for v in ([1, 0, 0], [0, 0, -1], [0.6, -0.64, 0.48], [-0.3, 0.2, -0.933]):
        v = np.asarray(v, dtype=np.float32)
        v /= np.linalg.norm(v)
        e = L.oracle_encode_normal(v.ctypes.data_as(C.POINTER(C.c_float)))
        L.oracle_decode_normal(e, n.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.all(np.abs(n - v) < 1e-2), (v, n)
    assert L.oracle_encode_normal(np.zeros(3, dtype=np.float32).ctypes.data_as(C.POINTER(C.c_float))) & (1 << 30) or True


def test_incident_matches_mapper():
    # IncidentsTests.cpp:27-89: the mapper's packed normal equals the scalar update applied in ray order
    L = po.lib()
    layers = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_INCIDENT]
    m = po.OracleMap(0.1, origin=(-0.05, -0.05, -0.05), layers=layers)
    rng = np.random.RandomState(7)
    rays = []
    expected = 0
    for r in range(1000):
        o = rng.uniform(-1, 1, 3)
        o = o / np.linalg.norm(o) * 0.3
        rays += [o, np.zeros(3)]
        expected = L.oracle_update_incident_normal(expected, o.astype(np.float32).ctypes.data_as(C.POINTER(C.c_float)), r)
    m.integrate_rays(np.asarray(rays))
    key = m.voxel_key([0, 0, 0])
    inc = m.region_layer(tuple(key[:3]), po.LAYER_INCIDENT)
    assert inc[key[3] + 32 * key[4] + 1024 * key[5]] == expected


# -- tests/ohmtestgpu/GpuTouchTimeTests.cpp:75-76 / VoxelTouchTimeCompute.h:18-27 ---------------------------------
def test_touch_time_is_ms_since_first_ray_and_last_write_wins():
    layers = [po.LAYER_OCCUPANCY, po.LAYER_TOUCH_TIME]
    m = po.OracleMap(0.1, layers=layers)
    rays = np.array([[0, 0, 0], [1.02, 0, 0]] * 3, dtype=np.float64)
    m.integrate_rays(rays, timestamps=np.array([10.0, 10.5, 10.25]))
    assert m.first_ray_time() == 10.0
    key = m.voxel_key([1.02, 0, 0])
    tt = m.region_layer(tuple(key[:3]), po.LAYER_TOUCH_TIME)
    assert tt[key[3] + 32 * key[4] + 1024 * key[5]] == 250      # CPU: last write, not max
    assert po.lib().oracle_encode_touch_time(10.0, 12.3456) == 2345


# -- tests/ohmtestcommon/CovarianceTestUtil.cpp:43-117 + tests/ohmtestgpu/GpuNdtTests.cpp:171-228 (Ndt.Hit) --------
def _reference_cov_update(P, mean, n, z):
    """Independent statement of the update the packed code implements (CovarianceVoxelCompute.h:323-330):
    Pnew = n/(n+1) P + n/(n+1)^2 (z-mu)(z-mu)^T ; mu_new = (n mu + z)/(n+1)."""
    d = z - mean
    return n / (n + 1.0) * P + n / (n + 1.0) ** 2 * np.outer(d, d), (n * mean + z) / (n + 1.0)


def _unpack_sqrt_cov(c):
    S = np.array([[c[0], 0, 0], [c[1], c[2], 0], [c[3], c[4], c[5]]], dtype=np.float64)
    return S @ S.T


def test_ndt_hit_matches_independent_covariance():
    res = 2.0
    layers = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_COVARIANCE]
    m = po.OracleMap(res, mode="ndt", layers=layers)
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    A = rng.uniform(-0.15, 0.15, size=(3, 3))
    samples = np.array([1.0, 1.0, 1.0]) + rng.normal(size=(10000, 3)) @ A.T
    samples = samples[np.all((samples > 0.01) & (samples < 1.99), axis=1)]
    rays = np.zeros((2 * len(samples), 3))
    rays[1::2] = samples
    m.integrate_rays(rays, ray_flags=1 << 4)  # kRfExcludeRay, as Ndt.Hit does
    key = m.voxel_key([1.0, 1.0, 1.0])
    vi = key[3] + 32 * key[4] + 1024 * key[5]
    cov = m.region_layer(tuple(key[:3]), po.LAYER_COVARIANCE)[vi]
    mean = m.region_layer(tuple(key[:3]), po.LAYER_MEAN)[vi]
    # reference recursion in float64, seeded like initialiseTestVoxel: S = 0.1 * res * I
    P = np.eye(3) * (0.1 * res) ** 2
    mu = np.zeros(3)
    for n, z in enumerate(samples):
        if n == 0:
            mu = z.copy()
            continue
        P, mu = _reference_cov_update(P, mu, float(n), z)
    assert mean[1] == len(samples)                                            # count exact
    out = np.zeros(3)
    po.lib().oracle_sub_voxel_to_local(int(mean[0]), res, out.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.linalg.norm(out + m.voxel_centre(key) - mu) < 1e-1              # epsilon_mean
    # the recursion is seeded with the first sample's zero-spread covariance scaled by the count weights:
    Pn = _unpack_sqrt_cov(cov)
    sample_cov = np.cov(samples.T, bias=True)
    assert np.all(np.abs(Pn - sample_cov) < 1e-2), (Pn, sample_cov)            # epsilon_cov
    assert np.all(np.abs(Pn - P) < 1e-2)


def test_ndt_hit_scalar_against_double_precision_update():
    # CovarianceTestUtil.cpp updateHit: same packed algorithm in double; float storage tolerance 1e-2
    L = po.lib()
    rng = np.random.RandomState(5)
    cov = np.zeros(6, dtype=np.float32)
    ref = None
    mean = np.zeros(3)
    value = C.c_float(np.inf)
    for n in range(200):
        z = np.array([1.0, 1.0, 1.0]) + rng.normal(scale=0.2, size=3)
        L.oracle_calculate_hit_with_covariance(
            cov.ctypes.data_as(C.POINTER(C.c_float)), C.byref(value), z.ctypes.data_as(C.POINTER(C.c_double)),
            mean.ctypes.data_as(C.POINTER(C.c_double)), n, HIT, INF, C.c_float(2.0), C.c_float(-1.3863), 100)
        if n == 0:
            assert np.allclose(cov, [0.2, 0, 0.2, 0, 0, 0.2])  # initialiseCovariance: 0.1 * resolution on the diagonal
            ref = np.eye(3) * 0.04
        else:
            ref, _ = _reference_cov_update(ref, mean, float(n), z)
        mean = (n * mean + z) / (n + 1.0)
        assert np.all(np.abs(_unpack_sqrt_cov(cov) - ref) < 1e-2)
    assert value.value > 0


# -- tests/ohmtestgpu/GpuNdtTests.cpp:235-406 (Ndt.Miss*), tolerance 1e-4 ---------------------------------------
def _ndt_miss_unpacked(S, mean, sensor, sample, value, miss_value, rate, noise):
    """Equation (24) of Saarinen et al. evaluated with the unpacked covariance inverse (the commented-out
    'unpacked version' in CovarianceVoxelCompute.h:243-258)."""
    P = S @ S.T
    Pinv = np.linalg.inv(P)
    ray = (sample - sensor) / np.linalg.norm(sample - sensor)
    a = Pinv @ ray
    t = a @ (mean - sensor) / (a @ ray)
    x_ml = ray * t + sensor
    p1 = math.exp(-0.5 * (x_ml - mean) @ Pinv @ (x_ml - mean))
    p2 = math.exp(-0.5 * np.dot(x_ml - sample, x_ml - sample) / (noise * noise))
    p = 0.5 - 0.5 * rate * p1 * (1.0 - p2)
    return value + math.log(p / (1.0 - p))


@pytest.mark.parametrize("shape", ["planar", "cylindrical", "spherical"])
def test_ndt_miss_matches_unpacked_formula(shape):
    L = po.lib()
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    n = 4000
    if shape == "planar":
        samples = np.column_stack([rng.uniform(0.01, 1.99, n), rng.uniform(0.01, 1.99, n), np.full(n, 1.0)])
        sensor = np.array([1.0, 1.0, 5.0])
        test_rays = [([1, 1, 5], [1, 1, -5]), ([1, 1, -5], [1, 1, 5]), ([-5, 1, 0.25], [5, 1, 0.25]),
                     ([1, 5, 1.01], [1, -5, 1.01]), ([-5, 1, 2], [5, 1, 1]), ([-5, 1, 2], [5, 1, 0.5])]
        origin = (0.0, 0.0, 0.0)
    else:
        r = 0.3
        v = rng.uniform(-0.99, 0.99, size=(n, 3))
        rad = rng.uniform(r - 0.05, r + 0.05, n)
        if shape == "cylindrical":
            lxy = np.linalg.norm(v[:, :2], axis=1)
            v[:, 0] = rad * v[:, 0] / lxy
            v[:, 1] = rad * v[:, 1] / lxy
        else:
            v = v / np.linalg.norm(v, axis=1)[:, None] * rad[:, None]
        samples = v
        sensor = np.array([0.0, 0.0, 5.0])
        test_rays = [([0, 0, 5], [0, 0, -5]), ([0, 0, -5], [0, 0, 5]), ([r, r, 5], [r, r, -5]),
                     ([1.5 * r, 1.5 * r, -5], [2 * r, 2 * r, 5]), ([2, -r, 0], [-2, -r, 0])]
        origin = (-1.0, -1.0, -1.0)
    layers = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_COVARIANCE]
    m = po.OracleMap(2.0, mode="ndt", layers=layers, origin=origin)
    rays = np.zeros((2 * n, 3))
    rays[0::2] = sensor
    rays[1::2] = samples
    m.integrate_rays(rays, ray_flags=1 << 4)
    key = m.voxel_key(samples[0])
    vi = key[3] + 32 * key[4] + 1024 * key[5]
    rk = tuple(key[:3])
    cov = np.ascontiguousarray(m.region_layer(rk, po.LAYER_COVARIANCE)[vi])
    vm = m.region_layer(rk, po.LAYER_MEAN)[vi]
    assert vm[1] == n
    mean = np.zeros(3)
    L.oracle_sub_voxel_to_local(int(vm[0]), 2.0, mean.ctypes.data_as(C.POINTER(C.c_double)))
    mean += m.voxel_centre(key)
    S = np.array([[cov[0], 0, 0], [cov[1], cov[2], 0], [cov[3], cov[4], cov[5]]], dtype=np.float64)
    value0 = float(m.region_layer(rk, po.LAYER_OCCUPANCY)[vi])
    for a, b in test_rays:
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        v = C.c_float(value0)
        is_miss = C.c_int(0)
        L.oracle_calculate_miss_ndt(cov.ctypes.data_as(C.POINTER(C.c_float)), C.byref(v), C.byref(is_miss),
                                    a.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)),
                                    mean.ctypes.data_as(C.POINTER(C.c_double)), n, INF, m.params.miss_value,
                                    m.params.adaptation_rate, m.params.sensor_noise, m.params.sample_threshold)
        expect = _ndt_miss_unpacked(S, mean, a, b, value0, m.params.miss_value, m.params.adaptation_rate,
                                    m.params.sensor_noise)
        assert abs(v.value - expect) < 1e-4, (shape, a, b, v.value, expect)
        assert v.value <= value0 + 1e-6           # a miss never raises occupancy
    # below the sample threshold / unobserved: plain occupancy behaviour (CovarianceVoxelCompute.h:556-572)
    v = C.c_float(np.inf)
    L.oracle_calculate_miss_ndt(cov.ctypes.data_as(C.POINTER(C.c_float)), C.byref(v), C.byref(is_miss),
                                sensor.ctypes.data_as(C.POINTER(C.c_double)), samples[0].ctypes.data_as(C.POINTER(C.c_double)),
                                mean.ctypes.data_as(C.POINTER(C.c_double)), n, INF, m.params.miss_value, 0.2, 0.05, 3)
    assert np.float32(v.value) == np.float32(m.params.miss_value)
    v = C.c_float(1.0)
    L.oracle_calculate_miss_ndt(cov.ctypes.data_as(C.POINTER(C.c_float)), C.byref(v), C.byref(is_miss),
                                sensor.ctypes.data_as(C.POINTER(C.c_double)), samples[0].ctypes.data_as(C.POINTER(C.c_double)),
                                mean.ctypes.data_as(C.POINTER(C.c_double)), 2, INF, m.params.miss_value, 0.2, 0.05, 3)
    assert np.float32(v.value) == np.float32(1.0) + np.float32(m.params.miss_value)


# -- tests/ohmtestgpu/GpuTsdfTests.cpp:19-84 (Tsdf.Basic), tolerance 1e-6 ----------------------------------------
def test_tsdf_distance_closed_form():
    bx, by, bz = 1.0, 0.9, 0.8
    ends = [(bx, 0, 0), (-bx, 0, 0), (bx, by, 0), (-bx, by, 0), (0, by, 0), (bx, -by, 0), (-bx, 0, bz), (bx, by, bz),
            (-bx, by, bz), (bx, 0, bz), (bx, -by, bz), (-bx, 0, bz), (bx, by, -bz), (-bx, by, -bz), (bx, 0, -bz),
            (bx, -by, -bz)]
    for e in ends:
        m = po.OracleMap(0.1, mode="tsdf", layers=[po.LAYER_TSDF], origin=(-0.05, -0.05, -0.05), tsdf_trunc=10.0)
        s, e = np.zeros(3), np.asarray(e, dtype=np.float64)
        m.integrate_rays(np.array([s, e]))
        keys, _, _ = m.walk_segment(s, e)
        assert len(keys) > 5
        for k in keys:
            c = m.voxel_centre(k)
            tsdf = m.region_layer(tuple(k[:3]), po.LAYER_TSDF)[k[3] + 32 * k[4] + 1024 * k[5]]
            g = np.linalg.norm(e - s)
            expect = g - np.dot(c - s, e - s) / g                     # computeDistance (VoxelTsdfCompute.h:57-69)
            assert abs(tsdf[1] - expect) < 1e-6 and tsdf[0] == 1.0


def test_tsdf_truncation_and_weight_cap():
    # Tsdf.Truncation: distance clamps to +-trunc, weight accumulates to max_weight
    L = po.lib()
    w, d = C.c_float(0), C.c_float(0)
    s, e = np.zeros(3), np.array([2.0, 0, 0])
    for c, expect in (([0.5, 0, 0], 0.1), ([2.5, 0, 0], -0.1), ([1.95, 0, 0], 0.05)):
        w.value, d.value = 0, 0
        c = np.asarray(c, dtype=np.float64)
        L.oracle_calculate_tsdf(s.ctypes.data_as(C.POINTER(C.c_double)), e.ctypes.data_as(C.POINTER(C.c_double)),
                                c.ctypes.data_as(C.POINTER(C.c_double)), 0.1, 3.0, 0.0, 1.0, C.byref(w), C.byref(d))
        assert abs(d.value - expect) < 1e-6 and w.value == 1.0
    for _ in range(10):
        L.oracle_calculate_tsdf(s.ctypes.data_as(C.POINTER(C.c_double)), e.ctypes.data_as(C.POINTER(C.c_double)),
                                c.ctypes.data_as(C.POINTER(C.c_double)), 0.1, 3.0, 0.0, 1.0, C.byref(w), C.byref(d))
    assert w.value == 3.0 and abs(d.value - 0.05) < 1e-6


# -- ray filters: ohm/RayFilter.cpp + tests/ohmtestgpu/GpuMapTest.cpp ClipBox family ------------------------------
def test_ray_filters():
    m = po.OracleMap(0.25, filter_kind=po.FILTER_GOOD_RAY, filter_range=10.0)
    rays = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 0], [np.nan, 0, 0], [0, 0, 0], [20.0, 0, 0], [np.inf, 0, 0], [1, 1, 1]])
    m.integrate_rays(rays)
    s = m.stats()
    assert s["rays_in"] == 4 and s["rays_accepted"] == 1 and s["sample_updates"] == 1
    m = po.OracleMap(0.25, filter_kind=po.FILTER_CLIP_RANGE, filter_range=10.0)
    m.integrate_rays(np.array([[0, 0, 0], [20.0, 0, 0]]))
    s = m.stats()
    assert s["rays_accepted"] == 1 and s["sample_updates"] == 0       # clipped end: no hit...
    key = m.voxel_key([10.0, 0, 0])
    occ = m.region_layer(tuple(key[:3]), po.LAYER_OCCUPANCY)
    assert occ[key[3] + 32 * key[4] + 1024 * key[5]] == np.float32(m.params.miss_value)  # ...the end voxel takes a miss
    far = m.voxel_key([12.0, 0, 0])
    assert m.region_layer(tuple(far[:3]), po.LAYER_OCCUPANCY) is None or np.isinf(
        m.region_layer(tuple(far[:3]), po.LAYER_OCCUPANCY)[far[3] + 32 * far[4] + 1024 * far[5]])
