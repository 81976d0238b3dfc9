#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the REFERENCE itself (ohm's own RayMapperOccupancy / Ndt / Tsdf and
walkSegmentKeys, compiled unmodified from /root/reference into oracle/_ref) on small seeded inputs.

    python tools/make_golden.py            # needs /root/reference (only present in the build container)

The fixtures pin the oracle (tests/test_golden.py, CPU) and the CUDA path (tests/test_gpu_golden.py) where the
reference is not available — /root/reference does not exist on the GPU box.  Every fixture holds its inputs (rays,
intensities, timestamps, the call's flags, the map parameters as JSON) and, per region, every layer the reference wrote.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def cases():
    """name -> dict(mode, resolution, params, passes=[(rays, intensities, timestamps, ray_flags)])"""
    from ohm_b200.lidar import cube_rays

    out = {}
    # BASELINE config 1: 10k rays into a 0.2 m map, a single region
    out["config1_occupancy"] = dict(mode="occupancy", resolution=0.2, params=dict(layers=[0, 1]),
                                    passes=[(cube_rays(10000), None, None, 0)])
    rng = np.random.RandomState(7)

    def sensor_rays(n, extent, origin=(0.05, 0.05, 0.05)):
        rays = np.empty((2 * n, 3))
        rays[0::2] = origin
        rays[1::2] = rng.uniform(-extent, extent, size=(n, 3))
        return rays

    # every sample layer, timestamps, two passes with different ray flags, saturation, odd origin
    n = 1500
    r1, r2 = sensor_rays(n, 5.0), sensor_rays(n, 5.0, origin=(0.6, -0.4, 0.3))
    out["occupancy_layers_flags"] = dict(
        mode="occupancy", resolution=0.25,
        params=dict(layers=[0, 1, 2, 3, 4], region_dim=(16, 16, 16), origin=(0.3, -0.7, 0.11), saturate_min=1,
                    min_value=-1.0),
        passes=[(r1, None, 100.0 + np.arange(n) * 1e-3, 0),
                (r2, None, 101.5 + np.arange(n) * 1e-3, (1 << 0) | (1 << 2)),   # end point as free, exclude origin
                (r1[::-1].copy(), None, 103.0 + np.arange(n) * 1e-3, 1 << 6)])  # exclude free
    # clip-range filter with bad rays
    bad = sensor_rays(600, 20.0)
    bad[7] = np.nan
    bad[100] = np.inf
    out["occupancy_clip_filter"] = dict(mode="occupancy", resolution=0.25,
                                        params=dict(region_dim=(16, 16, 16), filter_kind=2, filter_range=8.0),
                                        passes=[(bad, None, None, 0)])
    # NDT-TM: samples on planes so that voxels become Gaussians and are then traversed
    n = 3000
    pts = rng.uniform(-5, 5, size=(n, 3))
    pts[:1500, 2] = -1.0 + rng.normal(scale=0.02, size=1500)
    pts[1500:2200, 0] = 3.0 + rng.normal(scale=0.02, size=700)
    nd = np.empty((2 * n, 3))
    nd[0::2] = (0.05, 0.05, 0.05)
    nd[1::2] = pts
    inten = rng.uniform(0, 255, size=n).astype(np.float32)
    out["ndt_tm"] = dict(mode="ndt_tm", resolution=0.25, params=dict(region_dim=(16, 16, 16)),
                         passes=[(nd, inten, None, 0), (nd[::-1].copy(), inten[::-1].copy(), None, 0)])
    out["tsdf"] = dict(mode="tsdf", resolution=0.1, params=dict(region_dim=(16, 16, 16)),
                       passes=[(sensor_rays(1200, 2.5), None, None, 0)])
    return out


def main():
    from oracle import pyref

    if not pyref.available():
        raise SystemExit("oracle/_ref is not built: this script needs /root/reference")
    os.makedirs(OUT, exist_ok=True)
    for name, case in cases().items():
        m = pyref.ReferenceMap(case["resolution"], mode=case["mode"], **case["params"])
        arrays = {}
        for i, (rays, intensities, timestamps, flags) in enumerate(case["passes"]):
            m.integrate_rays(rays, intensities, timestamps, flags)
            arrays[f"pass{i}_rays"] = np.ascontiguousarray(rays, dtype=np.float64)
            if intensities is not None:
                arrays[f"pass{i}_intensities"] = np.asarray(intensities, dtype=np.float32)
            if timestamps is not None:
                arrays[f"pass{i}_timestamps"] = np.asarray(timestamps, dtype=np.float64)
            arrays[f"pass{i}_flags"] = np.asarray([flags], dtype=np.uint32)
        dump = m.dump()
        for key, layers in dump.items():
            for layer, data in layers.items():
                arrays["region_%d_%d_%d_layer%d" % (key + (layer,))] = np.ascontiguousarray(data)
        meta = dict(mode=case["mode"], resolution=case["resolution"], params=case["params"], passes=len(case["passes"]),
                    regions=len(dump), first_ray_time=m.first_ray_time(),
                    source="ohm::RayMapper* of csiro-robotics/ohm @ 4e2e769, compiled from /root/reference (oracle/_ref)")
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        m.close()
        print(name, len(dump), "regions", os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")
    # line walk: keys and enter / exit ranges of walkSegmentKeys for seeded rays (LineWalkTests' seed)
    m = pyref.ReferenceMap(0.25)
    rng = np.random.RandomState(1153297050 % 2 ** 32)
    starts = rng.uniform(-10, 10, size=(60, 3))
    ends = rng.uniform(-10, 10, size=(60, 3))
    walk = {"starts": starts, "ends": ends}
    for i in range(60):
        for flags in (0, 1, 2, 3):
            keys, enter, exit_ = m.walk_segment(starts[i], ends[i], flags)
            walk[f"keys_{i}_{flags}"] = np.asarray(keys, dtype=np.int32)
            walk[f"enter_{i}_{flags}"] = np.asarray(enter, dtype=np.float64)
            walk[f"exit_{i}_{flags}"] = np.asarray(exit_, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "linewalk.npz"), **walk)
    print("linewalk", os.path.getsize(os.path.join(OUT, "linewalk.npz")) // 1024, "KiB")
    m.close()


if __name__ == "__main__":
    main()
