#!/usr/bin/env python
"""Phase clocks of walkRegions (debug build only).

    nvcc ... -DOHMB200_PHASE_CLOCKS -o ohm_b200/libohmb200_dbg.so ohm_b200/csrc/ohmb200.cu
    OHMB200_LIB=ohm_b200/libohmb200_dbg.so OHMB200_GRAPHS=0 python tools/phase_clocks.py run > gpurun_out/phases.txt
    python tools/phase_clocks.py summarise gpurun_out/phases.txt

`run` integrates the config-2 sweep three times into a fresh map; every CTA of the instrumented kernel prints the SM
cycles thread 0 spent per phase.  `summarise` averages the last launch.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import numpy as np

    import ohm_b200
    from ohm_b200.lidar import LidarBox
    rays, _, _ = LidarBox(1).sweep()
    rays = np.ascontiguousarray(rays)
    gpu = ohm_b200.GpuMap(0.1, device_bytes=6 << 30)
    for _ in range(3):
        gpu.clear()
        gpu.integrate_rays(rays)
        gpu.sync_voxels()
        sys.stdout.flush()
        print("LAUNCH END", flush=True)
    gpu.close()


def summarise(path):
    launches, cur = [], []
    for line in open(path):
        if line.startswith("PH "):
            f = line.split()
            cur.append({f[i]: int(f[i + 1]) for i in range(2, len(f) - 1, 2)})
            cur[-1].setdefault("shared", 0)
            cur[-1].setdefault("foldshared", 0)
        elif line.startswith("LAUNCH END"):
            launches.append(cur)
            cur = []
    last = launches[-1]
    names = ["top", "zero", "build", "walk", "wait", "fold", "foldshared"]
    tot = {n: sum(c[n] for c in last) for n in names}
    items = sum(c["items"] for c in last)
    allc = sum(tot.values())
    print(f"{sum(c['shared'] for c in last)} of the items share their region with others")
    print(f"{len(last)} CTAs, {items} work items; cycles of thread 0 per phase, summed over CTAs")
    for n in names:
        print(f"  {n:6s} {tot[n]:12d}  {100.0 * tot[n] / allc:5.1f} %   {tot[n] / max(items, 1):9.0f} cycles/item")
    print(f"  per CTA: {allc / len(last):.0f} cycles")
    if "fbatches" in last[0]:
        nb = sum(c["fbatches"] for c in last)
        for n in ("fscan", "fload", "fapply"):
            print(f"  sole fold, per batch of thread 0: {n:7s} {sum(c[n] for c in last) / max(nb, 1):8.0f} cycles")
        print(f"  ({nb} batches)")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        summarise(sys.argv[2])
