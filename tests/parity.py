"""Helpers shared by the parity tests: run the same rays through the CUDA path (C ABI) and the CPU oracle."""
import numpy as np

import ohm_b200
from ohm_b200 import gpumap as gm
from oracle import pyoracle as po

FLOAT_LAYERS = {gm.LAYER_OCCUPANCY, gm.LAYER_TRAVERSAL, gm.LAYER_COVARIANCE, gm.LAYER_INTENSITY, gm.LAYER_TSDF}


def make_pair(resolution, mode="occupancy", device_bytes=1 << 30, **overrides):
    """(GpuMap, OracleMap) configured identically."""
    cls = {"occupancy": ohm_b200.GpuMap, "ndt": ohm_b200.GpuNdtMap, "ndt_tm": ohm_b200.GpuNdtMap,
           "tsdf": ohm_b200.GpuTsdfMap}[mode]
    kw = dict(overrides)
    if mode == "ndt_tm":
        gpu = cls(resolution, traversability=True, device_bytes=device_bytes, **kw)
    else:
        gpu = cls(resolution, device_bytes=device_bytes, **kw)
    # The oracle takes the *resolved* layer set and parameters of the GPU map.
    okw = dict(overrides)
    okw["layers"] = int(gpu.params.layers)
    okw["ndt_tm"] = int(gpu.params.ndt_tm)
    cpu = po.OracleMap(resolution, mode=mode, **okw)
    return gpu, cpu


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def compare_maps(gpu, cpu, exact_layers=None, tol_layers=None):
    """Asserts both maps hold the same regions and layers.  exact_layers are compared bit for bit; tol_layers is
    {layer: abs_tol}.  Returns a dict of summary numbers."""
    g = gpu.dump()
    c = cpu.dump()
    assert sorted(g.keys()) == sorted(c.keys()), (
        f"region sets differ: gpu-only {sorted(set(g) - set(c))[:5]} cpu-only {sorted(set(c) - set(g))[:5]}")
    layers = gpu.layers()
    assert layers == cpu.layers()
    if exact_layers is None:
        exact_layers = [l for l in layers if not (tol_layers and l in tol_layers)]
    summary = {"regions": len(g), "voxels_observed": 0}
    for key in sorted(g.keys()):
        for layer in layers:
            ga, ca = g[key][layer], c[key][layer]
            if layer in exact_layers:
                gb, cb = bits(ga), bits(ca)
                if not np.array_equal(gb, cb):
                    bad = np.argwhere(gb != cb)
                    i = tuple(bad[0])
                    raise AssertionError(
                        f"layer {gm.LAYER_NAMES[layer]} region {key}: {len(bad)} mismatching words; first at {i}: "
                        f"gpu={ga[i]!r} cpu={ca[i]!r}")
            elif tol_layers and layer in tol_layers:
                gf = np.nan_to_num(ga.astype(np.float64), posinf=1e30, neginf=-1e30)
                cf = np.nan_to_num(ca.astype(np.float64), posinf=1e30, neginf=-1e30)
                # tolerance = (rtol, atol): |gpu - cpu| <= atol + rtol * |cpu|
                rtol, atol = tol_layers[layer]
                err = np.abs(gf - cf) - (atol + rtol * np.abs(cf))
                worst = int(np.argmax(err))
                assert err.max() <= 0, (f"layer {gm.LAYER_NAMES[layer]} region {key}: gpu {gf.flat[worst]} vs cpu "
                                        f"{cf.flat[worst]} exceeds rtol {rtol} atol {atol}")
        if gm.LAYER_OCCUPANCY in g[key]:
            summary["voxels_observed"] += int(np.isfinite(g[key][gm.LAYER_OCCUPANCY]).sum())
    return summary


def integrate_both(gpu, cpu, rays, intensities=None, timestamps=None, ray_flags=0, batch=None):
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
    n = rays.shape[0] // 2
    step = n if batch is None else batch
    for s in range(0, n, step):
        e = min(n, s + step)
        ii = None if intensities is None else intensities[s:e]
        tt = None if timestamps is None else timestamps[s:e]
        gpu.integrate_rays(rays[2 * s:2 * e], ii, tt, ray_flags)
    # The CPU mapper is order-sensitive only through ray order, which batching preserves.
    cpu.integrate_rays(rays, intensities, timestamps, ray_flags)
    gpu.sync_voxels()


def check_counts(gpu, cpu):
    gs, cs = gpu.stats(), cpu.stats()
    assert gs["rays_accepted"] == cs["rays_accepted"]
    assert gs["voxel_visits"] == cs["voxel_visits"]
    assert gs["sample_updates"] == cs["sample_updates"]
    return gs
