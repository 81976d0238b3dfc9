#!/usr/bin/env python
"""bench.py — Mrays/s of the ray-integration hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (libohmb200.so through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference CPU mapper (oracle port) on host cores

Workload (N=1): BASELINE.json configs[1] — "GpuMap occupancy-only: 1 synthetic 64x2048 lidar sweep, 0.1 m voxels".
One step = one integrateRays pass of the whole sweep (131072 rays) into an EMPTY map (the map is cleared and L2 is
flushed between steps, outside the timed spans), so every step includes region creation and streams the full
map working set (> L2) from HBM.

  value   rays/s with the rays already resident in HBM (ohmb200_integrate_device), CUDA-event timed on the stream
          the kernels are launched on, max over ranks.
  e2e     the same step through the reference-facing call ohmb200_integrate with HOST (pinned) ray buffers:
          host->device copy of the rays, all kernels, device->host read of the step's counters, and the
          syncVoxels-equivalent download of every occupancy region chunk, all inside the timed span.
  N>1     weak scaling: a step is N consecutive sweeps of the moving sensor (one per GPU, BASELINE config 5's
          trajectory) integrated as ONE batch into ONE map whose regions are sharded over the GPUs
          (ohmb200_set_partition).  Rank r holds sweep r; one NCCL all-gather per step hands every GPU the whole
          batch, and each GPU applies the visits/samples that fall in the regions it owns.  The union of the N
          maps is bit-identical to one GPU (or the CPU mapper) integrating the same batch.  value = rays of all
          N sweeps / step time.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/sec (64-beam lidar, 0.1 m voxels)"
UNIT = "Mrays/s"
RESOLUTION = 0.1
WORKLOAD = "GpuMap occupancy-only: 1 synthetic 64x2048 lidar sweep, 0.1 m voxels"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons = [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    clocks.append(float(f[1]))
                    out["sm_max_mhz"] = float(f[2])
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            # "under load": the upper half of the samples (idle samples before/after the spans read low)
            clocks.sort()
            out["sm_mhz"] = float(np.median(clocks[len(clocks) // 2:]))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(clocks)
        return out


def sweep_rays(count=1):
    """The first `count` sweeps of the trajectory, as a list of (2n, 3) ray arrays (count == 1: the config-2 sweep)."""
    from ohm_b200.lidar import LidarBox
    box = LidarBox(count)
    sweeps = [np.ascontiguousarray(box.sweep()[0]) for _ in range(count)]
    return sweeps[0] if count == 1 else sweeps


def cpu_mapper():
    """(constructor, kind): the reference's own RayMapperOccupancy when oracle/_ref was built from /root/reference
    (it travels to the GPU box as a prebuilt .so), else the C port of it."""
    from oracle import pyoracle as po
    from oracle import pyref as pr
    po.lib()
    if pr.available(build=False):
        return pr.ReferenceMap, "reference"
    return po.OracleMap, "port"


def cpu_baseline(rays, reps):
    """ohm's CPU RayMapperOccupancy, single thread (the mapper is single-threaded by design,
    ohm/RayMapperOccupancy.h:25-27), on `reps` fresh-map passes of the same sweep."""
    ctor, kind = cpu_mapper()
    n = rays.shape[0] // 2
    times = []
    for _ in range(reps):
        m = ctor(RESOLUTION)
        t0 = time.perf_counter()
        m.integrate_rays(rays)
        times.append(time.perf_counter() - t0)
        m.close()
    t = float(np.median(times))
    what = ("ohm::RayMapperOccupancy built from the reference sources (oracle/_ref)" if kind == "reference"
            else "C port of ohm::RayMapperOccupancy (oracle/ohm_oracle.c)")
    return {"value": n / t / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"{reps} x full config-2 sweep ({n} rays) into a fresh map, median {t:.3f} s/sweep; {what}",
            "seconds_per_sweep": t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rays = sweep_rays()
    n = rays.shape[0] // 2
    ctor, kind = cpu_mapper()
    for _ in range(min(args.warmup, 1)):
        m = ctor(RESOLUTION)
        m.integrate_rays(rays)
        m.close()
    t_total = 0.0
    for _ in range(args.steps):
        m = ctor(RESOLUTION)
        t0 = time.perf_counter()
        m.integrate_rays(rays)
        t_total += time.perf_counter() - t0
        m.close()
    ms = 1e3 * t_total / args.steps
    value = n / (ms * 1e-3) / 1e6
    what = ("ohm::RayMapperOccupancy, the reference's own CPU mapper compiled unmodified from its sources (oracle/_ref)"
            if kind == "reference" else "C port of ohm::RayMapperOccupancy (oracle/ohm_oracle.c)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 walk / f32 log-odds", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n,
                   "note": f"{what}; 1 thread — the mapper is single-threaded by design (ohm/RayMapperOccupancy.h:25-27); "
                           "each step is the full sweep into a fresh map"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{args.steps} x full sweep of {n} rays"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import ohm_b200
    from ohm_b200 import gpumap as gm

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    # stdout carries the one JSON line and nothing else: whatever a library prints there (NCCL's version banner, for
    # one) is sent to stderr from here on; the line itself is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # One sweep per rank: the step's batch is the concatenation of the N sweeps, in sweep order.
    sweeps = [sweep_rays()] if world == 1 else sweep_rays(world)
    rays = np.concatenate(sweeps)
    n = rays.shape[0] // 2
    gpu = ohm_b200.GpuMap(RESOLUTION, device_bytes=int(args.device_gib * (1 << 30)), device=local_rank)
    if world > 1:
        gpu.set_partition(rank, world)
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)

    # Device-resident rays: every rank holds its own sweep; the gathered batch is the kernels' input.
    per = max(s.shape[0] // 2 for s in sweeps)
    pad = per * world
    mine = np.full((per * 2, 3), np.nan)  # NaN rays are rejected by the filter (padding to the longest sweep only)
    mine[:sweeps[rank].shape[0]] = sweeps[rank]
    d_slice = torch.from_numpy(mine).cuda()
    d_full = torch.empty((pad * 2, 3), dtype=torch.float64, device="cuda")
    if world == 1:
        d_full.copy_(d_slice)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # Pinned host buffers for the end-to-end arm.
    h_rays = torch.from_numpy(rays).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reset_map():
        gpu.clear()
        flush.fill_(1)
        torch.cuda.synchronize()

    # Multi-GPU: the all-gather of step k+1 is queued on its own stream as soon as step k's kernels are queued, so the
    # exchange runs beside the kernels (two gather buffers; every step still pays for exactly one all-gather).
    comm = torch.cuda.Stream()
    d_fulls = [d_full, torch.empty_like(d_full)]
    gathered = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    step_no = [0]

    def gather(k):
        with torch.cuda.stream(comm):
            comm.wait_event(consumed[k & 1])  # the batch that last read this buffer is done
            dist.all_gather_into_tensor(d_fulls[k & 1], d_slice)
            gathered[k & 1].record(comm)

    if world > 1:
        for e in consumed:
            e.record(stream)
        gather(0)

    def device_step():
        k = step_no[0]
        step_no[0] += 1
        with torch.cuda.stream(stream):
            if world > 1:
                stream.wait_event(gathered[k & 1])
            gpu.integrate_rays_device(d_fulls[k & 1].data_ptr() if world > 1 else d_full.data_ptr(), 2 * pad)
            if world > 1:
                consumed[k & 1].record(stream)
        if world > 1:
            gather(k + 1)

    def timed(fn, count, profile=False):
        total_ms = 0.0
        for _ in range(count):
            reset_map()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if profile:
                gpu.set_profiling(True)
            a.record(stream)
            fn()
            b.record(stream)
            b.synchronize()
            if profile:
                gpu.set_profiling(False)
            barrier()
            ms = torch.tensor([a.elapsed_time(b)], device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            total_ms += float(ms.item())
        return total_ms

    # ---- device-resident arm -----------------------------------------------------------------------------
    timed(device_step, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = gpu.stats()["kernel_launches"]
    gpu.kernel_times(reset=True)
    t_ms = timed(device_step, args.steps)
    launches = gpu.stats()["kernel_launches"] - launches0
    # A second, separately profiled pass gives per-kernel durations (events around every launch perturb the
    # whole-step time slightly, so it is not the pass `value` comes from).
    timed(device_step, max(3, min(args.steps, 10)), profile=True)
    ktimes = gpu.kernel_times(reset=True)
    reset_map()
    device_step()
    torch.cuda.synchronize()
    st = gpu.stats()
    visits_rank, samples_rank = st["voxel_visits"], st["sample_updates"]
    regions_rank = st["regions"]
    tot = torch.tensor([visits_rank, samples_rank, regions_rank], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    visits, samples, regions = (int(x) for x in tot.tolist())
    ms_per_step = t_ms / args.steps
    value = n / (ms_per_step * 1e-3) / 1e6

    # ---- end-to-end arm (host buffers through ohmb200_integrate + map download) ---------------------------
    occ_bytes = gpu.L.ohmb200_region_layer_bytes(gpu.h, gm.LAYER_OCCUPANCY)
    keys = gpu.region_keys()
    h_map = torch.empty(max(len(keys), 1) * occ_bytes, dtype=torch.uint8).pin_memory()
    keys_c = np.ascontiguousarray(keys, dtype=np.int16)
    stats_struct = gm.Stats()

    h_maps = [h_map, torch.empty_like(h_map).pin_memory()]

    def e2e_run(count):
        """`count` steps through the public API from host buffers, as a user drives it: ohmb200_integrate returns
        after queueing (H2D of the pinned rays on the copy stream + kernels), the counters are read back, and the
        step's result — every occupancy chunk — is snapshotted and downloaded asynchronously
        (ohmb200_read_regions_async) while the next step runs.  Every step integrates into a FRESH map
        (ohmb200_clear is inside the timed region), as the reference arm does.  Drained before the clock stops."""
        for i in range(count):
            gpu.clear()
            gpu.integrate_rays_ptr(h_rays.data_ptr(), 2 * n)
            gpu.L.ohmb200_get_stats(gpu.h, ctypes.byref(stats_struct))
            k = gpu.region_keys()
            gpu.region_layers_async(k, gm.LAYER_OCCUPANCY, h_maps[i & 1].data_ptr(), h_maps[i & 1].numel())
        gpu.download_wait()
        torch.cuda.synchronize()

    def timed_host(count):
        reset_map()
        barrier()
        t0 = time.perf_counter()
        e2e_run(count)
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        barrier()
        return float(dt.item())

    def serial_step_seconds(count):
        # the same calls without overlap (download waited for before the next step): latency of one step
        total = 0.0
        for _ in range(count):
            reset_map()
            barrier()
            t0 = time.perf_counter()
            e2e_run(1)
            total += time.perf_counter() - t0
        return total / count

    timed_host(max(1, min(args.warmup, 3)))
    e2e_s = timed_host(args.steps) / args.steps
    e2e_serial_s = serial_step_seconds(3)
    e2e_value = n / e2e_s / 1e6
    clocks = sampler.stop() if rank == 0 else None
    d2h = len(keys) * occ_bytes + ctypes.sizeof(gm.Stats) + keys_c.nbytes

    if rank == 0:
        peak, peak_src = measured_peak()
        # Algorithmic bytes (occupancy-only): 8 B per voxel visit (4 R + 4 W) + 44 B per ray.
        alg_bytes = 8 * visits + 44 * n
        dom = max(ktimes.items(), key=lambda kv: kv[1]["ms"]) if ktimes else ("walkRays", {"ms": 0, "launches": 1})
        dom_name, dom_t = dom
        dom_ms = dom_t["ms"] / max(dom_t["launches"], 1)
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 / max(world, 1) if dom_ms > 0 else 0.0
        traffic = ncu_traffic()
        cpu = cpu_baseline(sweeps[0], reps=args.cpu_reps) if args.cpu_reps > 0 and world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 walk / f32 log-odds", "data": "synthetic",
            "config": {
                "workload": WORKLOAD if world == 1 else
                f"{WORKLOAD} per GPU: {world} consecutive sweeps of the moving sensor as one batch into one region-sharded map",
                "rays_per_step": n, "voxel_visits_per_step": visits,
                "sample_updates_per_step": samples, "regions": regions, "resolution_m": RESOLUTION,
                "parallelism": "single GPU" if world == 1 else f"regions sharded over {world} GPUs (owner = (rx + 2 ry + 4 rz) mod {world}); rank r brings sweep r, 1 NCCL all-gather of the batch per step, queued one step ahead on its own stream",
                "l2": "map cleared + 512 MiB L2 flush between timed steps (outside the timed spans); the per-step map "
                      "working set (pending + occupancy tiles of every touched region) exceeds the 126 MB L2",
                "timing": "CUDA events on the launch stream per step, max over ranks, summed over steps",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_rays.numel() * 8),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                    "serial_ms_per_step": e2e_serial_s * 1e3,
                    "what": "per step: ohmb200_clear (fresh map) + ohmb200_integrate(pinned host rays) + counters read + "
                            "snapshot and asynchronous download of every occupancy chunk to pinned memory "
                            "(ohmb200_read_regions_async), which overlaps the next step; drained inside the timed "
                            "region.  serial_ms_per_step = the same calls with the download waited for each step"},
            "gpu_launches": int(launches),
            "kernels_ms_per_step": {k: v["ms"] / max(v["launches"], 1) for k, v in ktimes.items()},
            "roofline": {
                "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None,
                "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                "algorithmic_bytes_per_launch": alg_bytes // max(world, 1),
                "kernel_ms": dom_ms, "peak_source": peak_src,
                "model": "8 B x voxel visits + 44 B x rays (SURVEY §8d), per GPU",
            },
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ohmb200", choices=["ohmb200", "reference"])
    ap.add_argument("--device-gib", type=float, default=6.0, help="device bytes for the region slabs")
    ap.add_argument("--cpu-reps", type=int, default=8, help="oracle passes for cpu_baseline (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
