#!/usr/bin/env python
"""Measures BASELINE.json configs 3 and 4 (multi-sweep NDT and TSDF) on one B200 next to the reference CPU mapper,
and checks size-independent properties at full size.  Writes profiles/configs_<tag>.json.

    python tools/measure_configs.py --tag r1 [--ndt-sweeps 100] [--tsdf-sweeps 50] [--cpu-sweeps 3]

This is a reporting tool, not bench.py (whose contract is config 2).  Timing: CUDA events around the whole run of
sweeps with the rays already resident in HBM ("device"), and wall clock around ohmb200_integrate from host buffers
plus the final syncVoxels-style download of every layer ("e2e").
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def key_l1_visits(rays, resolution, include_end):
    """Independent count of the voxels a walk visits — 1 + |dx|+|dy|+|dz| per ray minus the excluded end voxel — from
    the oracle's key maths alone (ohm/MapCoord.h:37-93, epsilon snaps included); no walking, so it scales to full size."""
    from oracle import pyoracle as po
    m = po.OracleMap(resolution)
    n = m.count_walk_visits(rays, 0 if include_end else 2)
    m.close()
    return n


def run(mode, resolution, sweeps, cpu_sweeps, device_gib):
    import torch

    import ohm_b200
    from ohm_b200 import gpumap as gm
    from ohm_b200.lidar import LidarBox
    from oracle import pyref, pyoracle

    box = LidarBox(sweeps)
    data = [box.sweep() for _ in range(sweeps)]
    n_rays = sum(d[0].shape[0] // 2 for d in data)
    cls = ohm_b200.GpuNdtMap if mode == "ndt" else ohm_b200.GpuTsdfMap
    out = {"mode": mode, "resolution": resolution, "sweeps": sweeps, "rays": n_rays}

    # device-resident
    g = cls(resolution, device_bytes=int(device_gib * (1 << 30)))
    stream = torch.cuda.Stream()
    g.set_stream(stream.cuda_stream)
    d_rays = [torch.from_numpy(np.ascontiguousarray(d[0])).cuda() for d in data]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for t in d_rays:
        g.integrate_rays_device(t.data_ptr(), t.shape[0])
    b.record(stream)
    b.synchronize()
    g.sync_voxels()
    ms = a.elapsed_time(b)
    st = g.stats()
    out["device"] = {"ms_total": ms, "ms_per_sweep": ms / sweeps, "mrays_per_s": n_rays / ms / 1e3}
    out["stats"] = st
    # size-independent properties at full size
    expect_visits = sum(key_l1_visits(d[0], resolution, include_end=(mode == "tsdf")) for d in data)
    out["properties"] = {"voxel_visits_equals_key_l1": st["voxel_visits"] == expect_visits,
                         "voxel_visits": st["voxel_visits"], "expected": expect_visits}
    keys = g.region_keys()
    if mode == "ndt":
        total = 0
        occ_ok = True
        for i in range(0, len(keys), 512):
            mean = g.region_layers(keys[i:i + 512], gm.LAYER_MEAN)
            total += int(mean[..., 1].sum(dtype=np.int64))
            occ = g.region_layers(keys[i:i + 512], gm.LAYER_OCCUPANCY)
            fin = occ[np.isfinite(occ)]
            occ_ok = occ_ok and bool(fin.min() >= g.params.min_value - 1e-6) and bool(fin.max() <= g.params.max_value + 1e-6)
        out["properties"]["mean_counts_sum_equals_samples"] = total == st["sample_updates"]
        out["properties"]["occupancy_within_clamp"] = occ_ok
        out["properties"]["samples"] = st["sample_updates"]
    else:
        ok = True
        for i in range(0, len(keys), 256):
            t = g.region_layers(keys[i:i + 256], gm.LAYER_TSDF)
            w, d = t[..., 0], t[..., 1]
            ok = ok and bool(np.all(w >= 0)) and bool(np.all(w <= g.params.tsdf_max_weight))
            ok = ok and bool(np.all(np.abs(d) <= g.params.tsdf_trunc + 1e-7)) and bool(np.all(w == np.floor(w)))
        out["properties"]["tsdf_weight_and_distance_bounds"] = ok
    out["regions"] = int(len(keys))
    g.close()

    # per-kernel times (event pairs around every launch; serialises the side stream, so the sum exceeds ms_per_sweep)
    g = cls(resolution, device_bytes=int(device_gib * (1 << 30)))
    g.set_stream(stream.cuda_stream)
    g.set_profiling(True)
    for t in d_rays:
        g.integrate_rays_device(t.data_ptr(), t.shape[0])
    g.sync_voxels()
    out["kernels_ms_per_sweep"] = {k: v["ms"] / sweeps for k, v in g.kernel_times().items() if v["launches"]}
    g.close()

    # end to end: host rays through the reference-facing call + download of every layer
    g = cls(resolution, device_bytes=int(device_gib * (1 << 30)))
    t0 = time.perf_counter()
    for d in data:
        g.integrate_rays(d[0])
    g.sync_voxels()
    keys = g.region_keys()
    nbytes = 0
    for layer in g.layers():
        for i in range(0, len(keys), 1024):
            nbytes += g.region_layers(keys[i:i + 1024], layer).nbytes
    dt = time.perf_counter() - t0
    out["e2e"] = {"s_total": dt, "mrays_per_s": n_rays / dt / 1e6, "d2h_bytes": nbytes,
                  "what": "ohmb200_integrate per sweep from pageable host rays + final download of every layer chunk"}
    g.close()

    # reference CPU mapper on the first sweeps of the same trajectory
    if cpu_sweeps > 0:
        kind = "reference" if pyref.available(build=False) else "port"
        ctor = pyref.ReferenceMap if kind == "reference" else pyoracle.OracleMap
        kw = {} if kind == "reference" else ({"layers": [0, 1, 5]} if mode == "ndt" else {"layers": [8]})
        m = ctor(resolution, mode=mode, **kw)
        t0 = time.perf_counter()
        nr = 0
        for d in data[:cpu_sweeps]:
            m.integrate_rays(d[0])
            nr += d[0].shape[0] // 2
        dt = time.perf_counter() - t0
        out["cpu"] = {"kind": kind, "cores": 1, "sweeps": cpu_sweeps, "s_per_sweep": dt / cpu_sweeps,
                      "mrays_per_s": nr / dt / 1e6}
        m.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--ndt-sweeps", type=int, default=100)
    ap.add_argument("--tsdf-sweeps", type=int, default=50)
    ap.add_argument("--cpu-sweeps", type=int, default=3)
    args = ap.parse_args()
    results = {}
    if args.ndt_sweeps:
        results["config3_ndt"] = run("ndt", 0.1, args.ndt_sweeps, args.cpu_sweeps, 16.0)
        print(json.dumps(results["config3_ndt"]))
    if args.tsdf_sweeps:
        results["config4_tsdf"] = run("tsdf", 0.05, args.tsdf_sweeps, min(args.cpu_sweeps, 2), 40.0)
        print(json.dumps(results["config4_tsdf"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"configs_{args.tag}.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
