#!/usr/bin/env python
"""Measures the DRAM traffic of the dominant kernels with ncu and writes the file bench.py quotes as `roofline.traffic`:

    gpurun -- 'python tools/ncu_traffic.py'      ->  gpurun_out/traffic.json   (copy it to profiles/traffic.json)

One `ncu --set full` capture of one warm launch per kernel (configs 2, 3 and 4), dram__bytes_read.sum +
dram__bytes_write.sum per launch.  The file carries the hash of the CUDA sources it was measured on; bench.py prints
the number only when that hash matches the sources it is running (a stale capture is reported as null, never quoted).
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

RUNS = [("walkRegions", ["--config", "2"], "^walkRegions$"), ("walkRegionsNdt", ["--config", "3"], "^walkRegionsNdt$"),
        ("walkRegionsTsdf", ["--config", "4"], "^walkRegionsTsdf$")]
METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def main():
    out = {"source_hash": bench.source_hash(), "kernels": {},
           "how": "ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum, one warm launch "
                  "(tools/ncu_traffic.py), bytes per launch"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name, args, regex in RUNS:
        cmd = ["ncu", "--clock-control", "none", "--metrics", METRICS, "-k", f"regex:{regex}", "-s", "5", "-c", "1", "--csv",
               sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--cpu-reps", "0"] + args
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
        rows = [r for r in csv.reader(io.StringIO(res.stdout)) if r and not r[0].startswith("==")]
        hdr = next((r for r in rows if "Metric Name" in r), None)
        if not hdr:
            print(f"{name}: no ncu rows\n{res.stdout[-400:]}\n{res.stderr[-400:]}", file=sys.stderr)
            continue
        ni, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        vals = {}
        for r in rows[rows.index(hdr) + 1:]:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}.get(r[ui], 1)
            vals[r[ni]] = float(r[vi].replace(",", "")) * scale
        out["kernels"][name] = {
            "dram_bytes_per_launch": int(vals.get("dram__bytes_read.sum", 0) + vals.get("dram__bytes_write.sum", 0)),
            "dram_bytes_read": int(vals.get("dram__bytes_read.sum", 0)),
            "dram_bytes_write": int(vals.get("dram__bytes_write.sum", 0)),
            "kernel_us_under_ncu": vals.get("gpu__time_duration.sum")}
        print(name, out["kernels"][name])
    with open(os.path.join(ROOT, "gpurun_out", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
