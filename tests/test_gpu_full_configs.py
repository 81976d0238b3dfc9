"""BASELINE configs 3 and 4 at FULL size on the GPU against the reference itself: 100 sweeps of the moving sensor into
an NDT map at 0.1 m, 50 sweeps into a TSDF map at 0.05 m — the multi-sweep interaction (Gaussians establishing and
re-initialising over 100 sweeps, TSDF weights growing towards saturation) is where an order-dependence bug would hide.
The CPU mapper needs minutes for these, so the comparison is against digests of ohm's own maps (RayMapperNdt /
RayMapperTsdf compiled from /root/reference), generated once by tools/make_golden_full.py and committed under
tests/golden/: per region a BLAKE2b digest of every bit-exact layer, and for the NDT log-odds (tolerance 1e-5, stated in
DESIGN.md §2) the count of observed voxels (exact) and the sum / sum of squares of their values."""
import hashlib
import os

import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from ohm_b200.lidar import LidarBox

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(arr):
    return np.frombuffer(hashlib.blake2b(np.ascontiguousarray(arr).tobytes(), digest_size=8).digest(), dtype=np.uint64)[0]


def load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: generate it with tools/make_golden_full.py where /root/reference exists")
    return np.load(path)


def run_trajectory(gpu, sweeps):
    box = LidarBox(sweeps)
    rays_total = 0
    for _ in range(sweeps):
        rays, _, _ = box.sweep()
        gpu.integrate_rays(rays)
        rays_total += rays.shape[0] // 2
    gpu.sync_voxels()
    return rays_total


def region_chunks(gpu, keys, layer, chunk=256):
    for lo in range(0, len(keys), chunk):
        part = keys[lo:lo + chunk]
        data = gpu.region_layers(part, layer)
        for i in range(len(part)):
            yield lo + i, data[i]


def test_config3_ndt_100_sweeps_matches_the_reference_map(gpu):
    g = load("full_config3.npz")
    sweeps = int(g["sweeps"])
    m = ohm_b200.GpuNdtMap(float(g["resolution"]), device_bytes=16 << 30)
    assert run_trajectory(m, sweeps) == int(g["rays"])
    keys = m.region_keys()
    order = np.lexsort((keys[:, 0], keys[:, 1], keys[:, 2]))          # (z, y, x) — any fixed order will do
    keys = keys[order]
    gold_keys = g["keys"]
    gold_order = np.lexsort((gold_keys[:, 0], gold_keys[:, 1], gold_keys[:, 2]))
    assert np.array_equal(keys, gold_keys[gold_order]), "region sets differ"
    for layer in (gm.LAYER_MEAN, gm.LAYER_COVARIANCE):
        want = g[f"digest_{layer}"][gold_order]
        bad = [i for i, arr in region_chunks(m, keys, layer) if digest(arr) != want[i]]
        assert not bad, f"{gm.LAYER_NAMES[layer]}: {len(bad)} of {len(keys)} regions differ from the reference, first {keys[bad[0]]}"
    count, total, total_sq = g["occ_count"][gold_order], g["occ_sum"][gold_order], g["occ_sum_sq"][gold_order]
    for i, occ in region_chunks(m, keys, gm.LAYER_OCCUPANCY):
        v = occ.astype(np.float64)
        fin = v[np.isfinite(v)]
        assert len(fin) == count[i], f"region {keys[i]}: {len(fin)} observed voxels, the reference has {count[i]}"
        # every value within 1e-5 (1 + |v|) of the reference's  =>  the sums within 1e-5 (n + sum|v|) (and so on)
        slack = 1e-5 * (len(fin) + np.abs(fin).sum())
        assert abs(fin.sum() - total[i]) <= slack, (keys[i], fin.sum(), total[i])
        assert abs((fin * fin).sum() - total_sq[i]) <= 8.0 * slack, (keys[i], (fin * fin).sum(), total_sq[i])
    st = m.stats()
    assert st["rays_accepted"] == int(g["rays"]) and st["sample_updates"] == int(g["rays"])
    m.close()


def test_config4_tsdf_50_sweeps_matches_the_reference_map(gpu):
    g = load("full_config4.npz")
    sweeps = int(g["sweeps"])
    m = ohm_b200.GpuTsdfMap(float(g["resolution"]), device_bytes=40 << 30)
    assert run_trajectory(m, sweeps) == int(g["rays"])
    keys = m.region_keys()
    order = np.lexsort((keys[:, 0], keys[:, 1], keys[:, 2]))
    keys = keys[order]
    gold_keys = g["keys"]
    gold_order = np.lexsort((gold_keys[:, 0], gold_keys[:, 1], gold_keys[:, 2]))
    assert np.array_equal(keys, gold_keys[gold_order]), "region sets differ"
    want = g[f"digest_{gm.LAYER_TSDF}"][gold_order]
    bad = [i for i, arr in region_chunks(m, keys, gm.LAYER_TSDF) if digest(arr) != want[i]]
    assert not bad, f"tsdf: {len(bad)} of {len(keys)} regions differ from the reference, first {keys[bad[0]]}"
    m.close()
