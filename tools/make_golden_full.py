#!/usr/bin/env python
"""Generates tests/golden/full_config{3,4}.npz: digests of the REFERENCE's own maps (oracle/_ref: ohm::RayMapperNdt /
RayMapperTsdf compiled unmodified from /root/reference) after the FULL BASELINE configs 3 and 4 —

    config 3: GpuNdtMap's CPU twin, 100 sweeps of the moving sensor, 0.1 m voxels (12.7 M rays, ~110 s on one core)
    config 4: TSDF, 50 sweeps, 0.05 m voxels (6.5 M rays)

— so that the multi-sweep interaction (Gaussians re-initialising over 100 sweeps, TSDF weights saturating) is held to
the reference at full size on the GPU box, where neither /root/reference nor minutes of CPU are available
(tests/test_gpu_full_configs.py).  Per region: a 64-bit BLAKE2b digest of every bit-exact layer; for the NDT log-odds
(compared with a tolerance) the count of observed voxels and the float64 sum and sum of squares of their values.

    python tools/make_golden_full.py [--sweeps3 100] [--sweeps4 50]      # needs /root/reference (build container only)
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def digest(arr):
    return np.frombuffer(hashlib.blake2b(np.ascontiguousarray(arr).tobytes(), digest_size=8).digest(), dtype=np.uint64)[0]


def log_odds_summary(occ):
    occ = np.asarray(occ, dtype=np.float64)
    fin = occ[np.isfinite(occ)]
    return len(fin), float(fin.sum()), float((fin * fin).sum())


def summarise(dump, exact_layers, occupancy_layer=None):
    keys = sorted(dump)
    out = {"keys": np.asarray(keys, dtype=np.int16)}
    for layer in exact_layers:
        out[f"digest_{layer}"] = np.asarray([digest(dump[k][layer]) for k in keys], dtype=np.uint64)
    if occupancy_layer is not None:
        s = [log_odds_summary(dump[k][occupancy_layer]) for k in keys]
        out["occ_count"] = np.asarray([x[0] for x in s], dtype=np.int64)
        out["occ_sum"] = np.asarray([x[1] for x in s])
        out["occ_sum_sq"] = np.asarray([x[2] for x in s])
    return out


def run(mode, resolution, sweeps):
    from ohm_b200.lidar import LidarBox
    from oracle import pyref
    assert pyref.available(build=True), "oracle/_ref is needed (the reference itself)"
    m = pyref.ReferenceMap(resolution, mode=mode)
    box = LidarBox(sweeps)
    t0 = time.perf_counter()
    rays_total = 0
    for k in range(sweeps):
        rays, _, _ = box.sweep()
        m.integrate_rays(rays)
        rays_total += rays.shape[0] // 2
        if k % 10 == 9:
            print(f"  {mode}: sweep {k + 1}/{sweeps}, {time.perf_counter() - t0:.0f} s", flush=True)
    return m, rays_total, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps3", type=int, default=100)
    ap.add_argument("--sweeps4", type=int, default=50)
    args = ap.parse_args()
    if args.sweeps3:
        m, n, dt = run("ndt", 0.1, args.sweeps3)
        out = summarise(m.dump(), exact_layers=[1, 5], occupancy_layer=0)   # mean, covariance | log-odds
        np.savez_compressed(os.path.join(OUT, "full_config3.npz"), sweeps=args.sweeps3, rays=n, resolution=0.1,
                            cpu_seconds=dt, **out)
        print(f"config 3: {n} rays, {len(out['keys'])} regions, {dt:.0f} s on the reference CPU mapper")
        m.close()
    if args.sweeps4:
        m, n, dt = run("tsdf", 0.05, args.sweeps4)
        out = summarise(m.dump(), exact_layers=[8])
        np.savez_compressed(os.path.join(OUT, "full_config4.npz"), sweeps=args.sweeps4, rays=n, resolution=0.05,
                            cpu_seconds=dt, **out)
        print(f"config 4: {n} rays, {len(out['keys'])} regions, {dt:.0f} s on the reference CPU mapper")
        m.close()


if __name__ == "__main__":
    main()
