#!/usr/bin/env python
"""bench.py — Mrays/s of the ray-integration hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (libohmb200.so through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference CPU mapper (oracle/_ref) on host cores
    python bench.py --config 3|4 [--steps 100]                # BASELINE configs 3 (NDT) and 4 (TSDF), same JSON line
    python bench.py --batch 4096                              # config 2 fed in the reference app's batches

N = 1, default: BASELINE.json configs[1] — "GpuMap occupancy-only: 1 synthetic 64x2048 lidar sweep, 0.1 m voxels".
One step = one integrateRays pass of the whole sweep (131072 rays) into an EMPTY map (the map is cleared and L2 is
flushed between steps, outside the timed spans), so every step includes region creation and streams the full map
working set (> L2) from HBM.  Configs 3 and 4 are trajectories: step k = sweep k of the moving sensor into the map the
sweeps before it built (warm-up sweeps first), L2 flushed between steps.

  value   rays/s with the rays already resident in HBM, CUDA-event timed on the stream the kernels are launched on,
          max over ranks.
  e2e     the same step through the reference-facing call with HOST (pinned) ray buffers: host->device copy of the
          rays, all kernels, device->host read of the step's counters and the syncVoxels-equivalent download of the
          occupancy chunks, all inside the timed span.
  N > 1   weak scaling of the SAME workload (the driver computes efficiency from value(N) / value(1)): a step is N
          consecutive sweeps of the moving sensor, rank r brings sweep r, integrated into ONE map whose regions are sharded
          over the GPUs through the library's routed exchange (ohmb200_exchange_send / _integrate): every rank filters
          and cuts only its own sweep; segments and samples travel to the GPU that owns their region over NVLink peer
          memory; nothing is all-gathered.  The exchange is inside the timed region of both arms.  The same run then
          measures BASELINE configs[4] ("config5": GpuNdtMap, moving sensor, region-sharded) on the same GPUs, and ends
          with a parity gate: sums of the per-rank counters against the oracle's, and the union of the per-rank NDT maps
          of a small seeded batch against the CPU mapper.
"""
import argparse
import ctypes
import hashlib
import json
import os
import platform
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/sec (64-beam lidar, 0.1 m voxels)"
UNIT = "Mrays/s"

CONFIGS = {
    2: {"workload": "GpuMap occupancy-only: 1 synthetic 64x2048 lidar sweep, 0.1 m voxels", "mode": "occupancy",
        "resolution": 0.1, "trajectory": False, "dtype": "f64 walk / f32 log-odds",
        "model": "8 B x voxel visits + 44 B x rays (SURVEY §8d)"},
    3: {"workload": "GpuNdtMap (NdtMode::kOccupancy) with voxel-mean + covariance, sweeps of the moving sensor, 0.1 m voxels",
        "mode": "ndt", "resolution": 0.1, "trajectory": True, "dtype": "f64 walk + f64 NDT / f32 storage",
        "model": "(8 + 8 + 24) B x miss visits + (8 + 16 + 48) B x sample voxels + 44 B x rays (SURVEY §8d)"},
    4: {"workload": "GpuTsdfMap TSDF integration, 0.05 m voxels, sweeps of the moving sensor", "mode": "tsdf",
        "resolution": 0.05, "trajectory": True, "dtype": "f64 walk / f32 TSDF",
        "model": "16 B x voxel visits + 68 B x rays (SURVEY §8d)"},
}


def algorithmic_bytes(config, visits, sample_voxels, rays):
    if config == 3:
        return 40 * visits + 72 * sample_voxels + 44 * rays
    if config == 4:
        return 16 * visits + 68 * rays
    return 8 * visits + 44 * rays


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def source_hash():
    """Hash of the CUDA sources: profiles/traffic.json is only quoted for the build it was captured on."""
    h = hashlib.sha256()
    src = os.path.join(ROOT, "ohm_b200", "csrc")
    for name in sorted(os.listdir(src)):
        with open(os.path.join(src, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json, written by
    tools/ncu_traffic.py) — only when it was captured on exactly these sources; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if t.get("source_hash") != source_hash():
            return None, "profiles/traffic.json is from another build: not quoted"
        k = t.get("kernels", {}).get(kernel)
        return (k or {}).get("dram_bytes_per_launch"), t.get("how")
    except Exception:
        return None, "no ncu capture for this build"


def host_cpu():
    model = platform.processor() or "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return model, os.cpu_count() or 1


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


def trajectory(count):
    """The first `count` sweeps of the moving sensor: list of (rays (2n,3), intensities, timestamps)."""
    from ohm_b200.lidar import LidarBox
    box = LidarBox(count)
    return [tuple(np.ascontiguousarray(a) for a in box.sweep()) for _ in range(count)]


def cpu_mapper(mode="occupancy"):
    """(constructor, kind): the reference's own mapper when oracle/_ref was built from /root/reference (it travels to
    the GPU box as a prebuilt .so), else the C port of it."""
    from oracle import pyoracle as po
    from oracle import pyref as pr
    po.lib()
    if pr.available(build=False):
        return (lambda res: pr.ReferenceMap(res, mode=mode)), "reference"
    layers = {"occupancy": [0], "ndt": [0, 1, 5], "tsdf": [8]}[mode]
    return (lambda res: po.OracleMap(res, mode=mode, layers=layers)), "port"


MAPPER_NAMES = {"occupancy": "RayMapperOccupancy", "ndt": "RayMapperNdt", "tsdf": "RayMapperTsdf"}


def cpu_baseline(config, sweeps, reps):
    """ohm's CPU mapper, single thread (single-threaded by design, ohm/RayMapperOccupancy.h:25-27), on a bounded sample:
    config 2: `reps` fresh-map passes of the sweep; configs 3/4: the first `reps` sweeps of the trajectory into one map."""
    cfg = CONFIGS[config]
    ctor, kind = cpu_mapper(cfg["mode"])
    times, rays_done = [], 0
    if cfg["trajectory"]:
        m = ctor(cfg["resolution"])
        for rays, intens, ts in sweeps[:reps]:
            t0 = time.perf_counter()
            m.integrate_rays(rays)
            times.append(time.perf_counter() - t0)
            rays_done += rays.shape[0] // 2
        m.close()
        value = rays_done / sum(times) / 1e6
        sample = f"the first {len(times)} sweeps of the trajectory ({rays_done} rays) into one map, {sum(times):.2f} s"
    else:
        rays = sweeps[0][0]
        for _ in range(reps):
            m = ctor(cfg["resolution"])
            t0 = time.perf_counter()
            m.integrate_rays(rays)
            times.append(time.perf_counter() - t0)
            m.close()
        t = float(np.median(times))
        value = rays.shape[0] // 2 / t / 1e6
        sample = f"{reps} x full config-2 sweep ({rays.shape[0] // 2} rays) into a fresh map, median {t:.3f} s/sweep"
    name = MAPPER_NAMES[cfg["mode"]]
    what = (f"ohm::{name} built from the reference sources (oracle/_ref)" if kind == "reference"
            else f"C port of ohm::{name} (oracle/ohm_oracle.c)")
    model, cores = host_cpu()
    return {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": f"{sample}; {what}",
            "host_cpu": model, "host_cores": cores,
            "hypothetical_all_cores": {"value": value * cores, "unit": UNIT,
                                       "note": "1-core rate x host cores: an embarrassingly-parallel ceiling the "
                                               "reference cannot reach on one map (its mappers are single-threaded)"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    ctor, kind = cpu_mapper(cfg["mode"])
    steps = args.steps
    if cfg["trajectory"]:
        # a bounded sample: the mapper runs ~1 s per sweep
        steps = min(args.steps, 12)
        sweeps = trajectory(min(args.warmup, 1) + steps)
    else:
        sweeps = trajectory(1)
    n_total, t_total = 0, 0.0
    if cfg["trajectory"]:
        m = ctor(cfg["resolution"])
        for k, (rays, _, _) in enumerate(sweeps):
            t0 = time.perf_counter()
            m.integrate_rays(rays)
            dt = time.perf_counter() - t0
            if k >= min(args.warmup, 1):
                t_total += dt
                n_total += rays.shape[0] // 2
        m.close()
    else:
        rays = sweeps[0][0]
        for k in range(min(args.warmup, 1) + steps):
            m = ctor(cfg["resolution"])
            t0 = time.perf_counter()
            m.integrate_rays(rays)
            dt = time.perf_counter() - t0
            m.close()
            if k >= min(args.warmup, 1):
                t_total += dt
                n_total += rays.shape[0] // 2
    ms = 1e3 * t_total / steps
    value = n_total / t_total / 1e6
    name = MAPPER_NAMES[cfg["mode"]]
    what = (f"ohm::{name}, the reference's own CPU mapper compiled unmodified from its sources (oracle/_ref)"
            if kind == "reference" else f"C port of ohm::{name} (oracle/ohm_oracle.c)")
    model, cores = host_cpu()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": cfg["workload"], "rays_per_step": n_total // steps,
                   "note": f"{what}; 1 thread — the mapper is single-threaded by design (ohm/RayMapperOccupancy.h:25-27); "
                           + ("each step is one sweep of the trajectory into the same map" if cfg["trajectory"]
                              else "each step is the full sweep into a fresh map")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "host_cpu": model, "host_cores": cores,
                         "sample": f"{steps} x sweep, {n_total} rays"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def own_sweep_index(step, rank, world):
    """Sweep of the trajectory rank `rank` brings in step `step`: the steps' shares, in rank order, are the trajectory."""
    return step * world + rank


def compare_union(dumps, ref, world, log_odds_rtol=1e-5):
    """The parity gate's map check: `dumps[r]` = {region key: {layer: array}} of rank r, `ref` the CPU mapper's map.
    No region on two ranks, every region on its owner, the union's region set equal to the reference's; every layer but
    the NDT log-odds bit-identical, the log-odds within |d| <= rtol (1 + |v|)."""
    from ohm_b200 import _lib, gpumap as gm
    owner_of = _lib.load().ohmb200_region_owner
    union, dup, misplaced = {}, 0, 0
    for r, d in enumerate(dumps):
        for key, layers in d.items():
            dup += key in union
            misplaced += owner_of((ctypes.c_int16 * 3)(*key), world) != r
            union[key] = layers
    sets_ok = sorted(union) == sorted(ref) and dup == 0 and misplaced == 0
    exact, worst = sets_ok, 0.0
    if sets_ok:
        for key, layers in ref.items():
            for layer, arr in layers.items():
                got = union[key][layer]
                if layer == gm.LAYER_OCCUPANCY:
                    a64 = np.nan_to_num(np.asarray(got).astype(np.float64), posinf=1e30)
                    b64 = np.nan_to_num(np.asarray(arr).astype(np.float64), posinf=1e30)
                    worst = max(worst, float((np.abs(a64 - b64) - log_odds_rtol * np.abs(b64)).max()))
                else:
                    exact = exact and np.array_equal(np.ascontiguousarray(got).view(np.uint8),
                                                     np.ascontiguousarray(arr).view(np.uint8))
    return {"regions": len(ref), "region_sets_equal_no_duplicates_every_region_on_its_owner": bool(sets_ok),
            "other_layers_bit_exact": bool(exact), "log_odds_worst_abs_err_beyond_rtol": worst,
            "ok": bool(sets_ok and exact and worst <= log_odds_rtol)}


def make_map(mode, resolution, device_bytes, device):
    import ohm_b200
    cls = {"occupancy": ohm_b200.GpuMap, "ndt": ohm_b200.GpuNdtMap, "tsdf": ohm_b200.GpuTsdfMap}[mode]
    return cls(resolution, device_bytes=device_bytes, device=device)


def run_single(args, torch):
    """N = 1: configs 2 (default), 3, 4."""
    from ohm_b200 import gpumap as gm

    cfg = CONFIGS[args.config]
    traj = cfg["trajectory"]
    n_sweeps = (args.warmup + args.steps) if traj else 1
    sweeps = trajectory(n_sweeps)
    device_gib = args.device_gib if args.device_gib else {2: 6.0, 3: 16.0, 4: 40.0}[args.config]
    gpu = make_map(cfg["mode"], cfg["resolution"], int(device_gib * (1 << 30)), 0)
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)
    d_sweeps = [torch.from_numpy(s[0]).cuda() for s in sweeps]
    h_sweeps = [torch.from_numpy(s[0]).pin_memory() for s in sweeps]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    batch = args.batch if args.batch > 0 else None

    exchange1 = args.pipeline == "exchange"
    if exchange1:
        # A/B: the routed-exchange pipeline with world = 1 (every record's owner is this GPU)
        gm.open_exchange([gpu], max(s[0].shape[0] // 2 for s in sweeps))

    def integrate_device(k):
        t = d_sweeps[k if traj else 0]
        n = t.shape[0] // 2
        if exchange1:
            gpu.exchange_send_device(t.data_ptr(), 2 * n)
            gpu.exchange_integrate()
            return
        if batch is None:
            gpu.integrate_rays_device(t.data_ptr(), 2 * n)
        else:
            for lo in range(0, n, batch):
                hi = min(n, lo + batch)
                gpu.integrate_rays_device(t.data_ptr() + lo * 48, 2 * (hi - lo))

    def between_steps():
        if not traj:
            gpu.clear()
        flush.fill_(1)
        torch.cuda.synchronize()

    def run_steps(first, count, profile=False):
        total = 0.0
        for k in range(first, first + count):
            between_steps()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if profile:
                gpu.set_profiling(True)
            a.record(stream)
            with torch.cuda.stream(stream):
                integrate_device(k)
            b.record(stream)
            b.synchronize()
            if profile:
                gpu.set_profiling(False)
            total += a.elapsed_time(b)
        return total

    def rays_of(first, count):
        return sum(sweeps[k if traj else 0][0].shape[0] // 2 for k in range(first, first + count))

    # ---- device-resident arm ---------------------------------------------------------------------------------
    run_steps(0, args.warmup)
    sampler = ClockSampler(0)
    sampler.start()
    st0 = gpu.stats()
    t_ms = run_steps(args.warmup, args.steps)
    st1 = gpu.stats()
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    n_timed = rays_of(args.warmup, args.steps)
    ms_per_step = t_ms / args.steps
    value = n_timed / (t_ms * 1e-3) / 1e6
    if traj:
        visits = (st1["voxel_visits"] - st0["voxel_visits"]) / args.steps
        samples = (st1["sample_updates"] - st0["sample_updates"]) / args.steps
        sample_voxels = (st1["sample_voxels"] - st0["sample_voxels"]) / args.steps
    # Per-kernel durations from a separately profiled pass (events around every launch serialise the two streams, so
    # it is not the pass `value` comes from).  Trajectories replay their last timed sweeps into a copy of the state
    # they met: a fresh map fed the same sweeps again.
    gpu.kernel_times(reset=True)
    prof_steps = max(3, min(args.steps, 10))
    if traj:
        gpu.clear()
        for k in range(args.warmup + args.steps - prof_steps):
            integrate_device(k)
        torch.cuda.synchronize()
        gpu.kernel_times(reset=True)
        run_steps(args.warmup + args.steps - prof_steps, prof_steps, profile=True)
    else:
        run_steps(0, prof_steps, profile=True)
    ktimes = gpu.kernel_times(reset=True)
    if not traj:
        between_steps()
        integrate_device(0)
        torch.cuda.synchronize()
        st = gpu.stats()
        visits, samples, sample_voxels = st["voxel_visits"], st["sample_updates"], st["sample_voxels"]
    regions = gpu.stats()["regions"]

    # ---- end-to-end arm --------------------------------------------------------------------------------------
    layer = gm.LAYER_TSDF if cfg["mode"] == "tsdf" else gm.LAYER_OCCUPANCY
    chunk = gpu.L.ohmb200_region_layer_bytes(gpu.h, layer)
    stats_struct = gm.Stats()

    def integrate_host(k):
        t = h_sweeps[k if traj else 0]
        n = t.shape[0] // 2
        if exchange1:
            gpu.exchange_send_ptr(t.data_ptr(), 2 * n)
            gpu.exchange_integrate()
            return
        if batch is None:
            gpu.integrate_rays_ptr(t.data_ptr(), 2 * n)
        else:
            for lo in range(0, n, batch):
                hi = min(n, lo + batch)
                gpu.integrate_rays_ptr(t.data_ptr() + lo * 48, 2 * (hi - lo))

    if not traj:
        keys = gpu.region_keys()
        h_maps = [torch.empty(max(len(keys), 1) * chunk, dtype=torch.uint8).pin_memory() for _ in range(2)]

        def e2e_run(count):
            """`count` steps as a user drives them: fresh map, ohmb200_integrate from pinned host rays, counters read
            back, and every occupancy chunk snapshotted and downloaded asynchronously while the next step runs;
            drained before the clock stops."""
            for i in range(count):
                gpu.clear()
                integrate_host(0)
                gpu.L.ohmb200_get_stats(gpu.h, ctypes.byref(stats_struct))
                k = gpu.region_keys()
                gpu.region_layers_async(k, layer, h_maps[i & 1].data_ptr(), h_maps[i & 1].numel())
            gpu.download_wait()
            torch.cuda.synchronize()

        def timed_e2e(count):
            between_steps()
            t0 = time.perf_counter()
            e2e_run(count)
            return time.perf_counter() - t0

        timed_e2e(4)  # (two input staging buffers: the batch graphs are recorded in steps 3 and 4)
        e2e_s = timed_e2e(args.steps) / args.steps
        serial = 0.0
        for _ in range(3):
            serial += timed_e2e(1)
        e2e_extra = {"serial_ms_per_step": serial / 3 * 1e3}
        e2e_rays = rays_of(0, 1)
        h2d = int(h_sweeps[0].numel() * 8)
        d2h = int(len(keys) * chunk + ctypes.sizeof(gm.Stats) + keys.nbytes)
        e2e_what = ("per step: ohmb200_clear (fresh map) + ohmb200_integrate(pinned host rays) + counters read + snapshot "
                    "and asynchronous download of every occupancy chunk to pinned memory (ohmb200_read_regions_async), "
                    "which overlaps the next step; drained inside the timed region.  serial_ms_per_step = the same calls "
                    "with the download waited for each step")
    else:
        # a trajectory: the sweeps of the timed steps again, from host buffers, into a map that holds the warm-up
        # sweeps; counters read every step; the whole map's layer downloaded once at the end (syncVoxels at the end of a
        # run, ohmapp/OhmAppGpu.cpp:262-268), inside the timed region.
        h_map = torch.empty(max(regions, 1) * chunk, dtype=torch.uint8).pin_memory()

        def e2e_pass():
            gpu.clear()
            for k in range(args.warmup):
                integrate_host(k)
            gpu.sync_voxels()
            between_steps()
            t0 = time.perf_counter()
            for k in range(args.warmup, args.warmup + args.steps):
                integrate_host(k)
                gpu.L.ohmb200_get_stats(gpu.h, ctypes.byref(stats_struct))
            keys = gpu.region_keys()
            for lo in range(0, len(keys), 1024):
                part = keys[lo:lo + 1024]
                gpu.region_layers_async(part, layer, h_map.data_ptr() + lo * chunk, len(part) * chunk)
            gpu.download_wait()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / args.steps, keys

        e2e_pass()               # untimed: staging buffers allocated, pages touched
        e2e_s, keys = e2e_pass()
        e2e_extra = {}
        e2e_rays = n_timed / args.steps
        h2d = int(sum(h_sweeps[k].numel() for k in range(args.warmup, args.warmup + args.steps)) * 8 / args.steps)
        d2h = int((len(keys) * chunk + keys.nbytes) / args.steps + ctypes.sizeof(gm.Stats))
        e2e_what = ("per step: ohmb200_integrate(pinned host rays of the sweep) + counters read; after the last step the "
                    f"{gm.LAYER_NAMES[layer]} chunk of every region is downloaded (syncVoxels at the end of a run), inside "
                    "the timed region and amortised over the steps")
    e2e_value = e2e_rays / e2e_s / 1e6
    clocks = sampler.stop()

    peak, peak_src = measured_peak()
    rays_per_step = n_timed / args.steps
    alg_bytes = algorithmic_bytes(args.config, visits, sample_voxels, rays_per_step)
    per_kernel = {k: v["ms"] / prof_steps for k, v in ktimes.items()}  # per step (a step may launch a kernel many times)
    dom_name = max(per_kernel, key=per_kernel.get) if per_kernel else "walkRegions"
    dom_ms = per_kernel.get(dom_name, 0.0)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_how = ncu_traffic({3: "walkRegionsNdt", 4: "walkRegionsTsdf"}.get(args.config, "walkRegions") if dom_name == "walkRegions" else dom_name)
    cpu = cpu_baseline(args.config, sweeps, args.cpu_reps) if args.cpu_reps > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic",
        "config": {
            "workload": cfg["workload"] + (f" ({args.warmup} warm-up + {args.steps} timed sweeps of the trajectory)" if traj else "")
            + (f", fed in batches of {batch} rays (ohmapp/OhmAppCpu.h:52)" if batch else ""),
            "rays_per_step": rays_per_step, "voxel_visits_per_step": visits, "sample_updates_per_step": samples,
            "sample_voxels_per_step": sample_voxels, "regions": regions, "resolution_m": cfg["resolution"],
            "parallelism": "single GPU" + (" (routed-exchange pipeline, world = 1)" if exchange1 else ""),
            "graphs": ("batches are replayed as CUDA graphs (same buffers and size every step)" if not traj and batch is None
                       else "no graph replay: every batch differs in buffer or size"),
            "l2": "512 MiB L2 flush between timed steps (outside the timed spans)" + ("" if traj else "; map cleared too")
                  + "; the per-step map working set exceeds the 126 MB L2",
            "timing": "CUDA events on the launch stream per step, summed over steps",
        },
        "e2e": dict({"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "ms_per_step": e2e_s * 1e3, "what": e2e_what}, **e2e_extra),
        "gpu_launches": int(launches),
        "kernels_ms_per_step": per_kernel,
        "roofline": {
            "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_how,
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms, "peak_source": peak_src,
            "model": cfg["model"],
            "whole_step": {"achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak if peak else None},
        },
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    gpu.close()
    return line


def run_sharded(args, torch, dist, rank, local_rank, world, gloo):
    """N > 1: the routed exchange.  Primary arm: config 2's workload, one sweep per GPU and step (weak scaling);
    then BASELINE configs[4] (NDT trajectory, region-sharded) and the parity gate."""
    import ohm_b200
    from ohm_b200 import gpumap as gm
    from ohm_b200.lidar import RAYS_PER_SWEEP
    from oracle import pyoracle as po

    cfg = CONFIGS[2]
    ndt_steps, ndt_warm = args.ndt_steps, 2
    n_traj = max(world, world * (ndt_warm + ndt_steps))
    sweeps = trajectory(n_traj)                      # every rank draws the whole noise stream, keeps its own sweeps
    per = RAYS_PER_SWEEP
    dev_bytes = int((args.device_gib or 6.0) * (1 << 30))

    def connect(m, per_rank):
        handle = m.exchange_open(rank, world, per_rank)
        handles = [None] * world
        dist.all_gather_object(handles, handle, group=gloo)
        m.exchange_connect(handles)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    gpu = make_map("occupancy", cfg["resolution"], dev_bytes, local_rank)
    connect(gpu, per)
    stream = torch.cuda.Stream()
    gpu.set_stream(stream.cuda_stream)
    mine = sweeps[rank][0]
    n_mine = mine.shape[0] // 2
    n_step = sum(sweeps[r][0].shape[0] // 2 for r in range(world))
    d_mine = torch.from_numpy(mine).cuda()
    h_mine = torch.from_numpy(mine).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def reset_map():
        gpu.clear()
        flush.fill_(1)
        torch.cuda.synchronize()

    def device_step():
        with torch.cuda.stream(stream):
            gpu.exchange_send_device(d_mine.data_ptr(), 2 * n_mine)
            gpu.exchange_integrate()

    def timed(count, profile=False):
        """`count` steps queued back to back — no host wait between them: each step is fresh map (ohmb200_clear, queued) + L2
        flush + a DEVICE-side barrier between the ranks (ohmb200_exchange_barrier), then the timed span [send + integrate]
        between two CUDA events on the launch stream.  The ranks enter every span together (the device barrier) and no
        rank's span waits for a host: what is timed is device time.  Per step the max over ranks; summed over steps."""
        barrier()
        spans = []
        if profile:
            gpu.set_profiling(True)
        for _ in range(count):
            with torch.cuda.stream(stream):
                gpu.clear()
                flush.fill_(1)
                gpu.exchange_barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            device_step()
            b.record(stream)
            spans.append((a, b))
        torch.cuda.synchronize()
        if profile:
            gpu.set_profiling(False)
        barrier()
        ms = torch.tensor([a.elapsed_time(b) for a, b in spans], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.sum().item())

    warmup = max(args.warmup, 3)  # the step graphs (one per inbox parity) are recorded in steps 2 and 3
    timed(warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = gpu.stats()["kernel_launches"]
    t_ms = timed(args.steps)
    launches = gpu.stats()["kernel_launches"] - launches0
    gpu.kernel_times(reset=True)
    prof_steps = max(3, min(args.steps, 10))
    timed(prof_steps, profile=True)
    ktimes = {k: v["ms"] / prof_steps for k, v in gpu.kernel_times(reset=True).items()}
    reset_map()
    barrier()
    device_step()
    torch.cuda.synchronize()
    st = gpu.stats()
    seg_out, smp_out = gpu.exchange_last_counts(world)
    # NVLink bytes this rank wrote in the step: 32 B per segment record and 16 B per sample record (occupancy-only map) whose
    # owner is another GPU, + the 64-byte walk constants of every own ray to each of the other GPUs (copy engines)
    nvlink_bytes = (32 * sum(c for r, c in enumerate(seg_out) if r != rank) + 16 * sum(c for r, c in enumerate(smp_out) if r != rank)
                    + 64 * n_mine * (world - 1))
    mine_counts = [st["rays_accepted"], st["voxel_visits"], st["sample_updates"], st["regions"], nvlink_bytes,
                   sum(seg_out), st["owned_visits"]]
    all_counts = [None] * world
    dist.all_gather_object(all_counts, mine_counts, group=gloo)
    all_ktimes = [None] * world
    dist.all_gather_object(all_ktimes, ktimes, group=gloo)
    accepted, visits, samples, regions = (sum(c[i] for c in all_counts) for i in range(4))
    nvlink_max = max(c[4] for c in all_counts)
    segments_total = sum(c[5] for c in all_counts)
    ms_per_step = t_ms / args.steps
    value = n_step / (ms_per_step * 1e-3) / 1e6

    # ---- end-to-end arm: host rays of the own sweep -> exchange -> own regions' occupancy chunks on the host -----
    occ_bytes = gpu.L.ohmb200_region_layer_bytes(gpu.h, gm.LAYER_OCCUPANCY)
    keys = gpu.region_keys()
    h_maps = [torch.empty(max(len(keys), 1) * occ_bytes * 2, dtype=torch.uint8).pin_memory() for _ in range(2)]
    stats_struct = gm.Stats()

    def e2e_run(count):
        for i in range(count):
            gpu.clear()
            gpu.exchange_send_ptr(h_mine.data_ptr(), 2 * n_mine)
            gpu.exchange_integrate()
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
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        barrier()
        return float(dt.item())

    timed_host(4)  # (two staging buffers x two inbox parities: the step graphs are recorded in steps 3 and 4)
    e2e_s = timed_host(args.steps) / args.steps
    e2e_value = n_step / e2e_s / 1e6
    clocks = sampler.stop() if rank == 0 else None
    d2h = len(keys) * occ_bytes + ctypes.sizeof(gm.Stats) + keys.nbytes
    gpu.close()
    barrier()

    # ---- BASELINE configs[4]: GpuNdtMap, moving sensor, regions sharded over the GPUs --------------------------------
    ndt = make_map("ndt", 0.1, dev_bytes, local_rank)
    connect(ndt, per)
    ndt.set_stream(stream.cuda_stream)
    own = [sweeps[own_sweep_index(k, rank, world)][0] for k in range(ndt_warm + ndt_steps)]
    d_own = [torch.from_numpy(s).cuda() for s in own]
    ndt_rays = 0
    ndt_spans = []
    barrier()
    for k in range(ndt_warm + ndt_steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)
            ndt.exchange_barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        with torch.cuda.stream(stream):
            ndt.exchange_send_device(d_own[k].data_ptr(), d_own[k].shape[0])
            ndt.exchange_integrate()
        b.record(stream)
        ndt_spans.append((a, b))
        if k >= ndt_warm:
            ndt_rays += sum(sweeps[k * world + r][0].shape[0] // 2 for r in range(world))
    torch.cuda.synchronize()
    barrier()
    ms = torch.tensor([a.elapsed_time(b) for a, b in ndt_spans[ndt_warm:]], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ndt_ms = float(ms.sum().item())
    ndt.sync_voxels()
    nst = ndt.stats()
    ndt_counts = [None] * world
    dist.all_gather_object(ndt_counts, [nst["rays_accepted"], nst["voxel_visits"], nst["sample_updates"], nst["regions"]],
                           group=gloo)
    ndt.close()
    barrier()
    # the same sweeps on ONE of these GPUs (plain single-GPU path), for the speed-up of the sharded run
    single_ms = None
    if rank == 0:
        one = make_map("ndt", 0.1, dev_bytes, local_rank)
        one.set_stream(stream.cuda_stream)
        d_all = [torch.from_numpy(sweeps[k][0]).cuda() for k in range(world * (ndt_warm + ndt_steps))]
        with torch.cuda.stream(stream):
            for k in range(world * ndt_warm):
                one.integrate_rays_device(d_all[k].data_ptr(), d_all[k].shape[0])
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        with torch.cuda.stream(stream):
            for k in range(world * ndt_warm, world * (ndt_warm + ndt_steps)):
                one.integrate_rays_device(d_all[k].data_ptr(), d_all[k].shape[0])
        b.record(stream)
        b.synchronize()
        single_ms = a.elapsed_time(b)
        one_st = one.stats()
        one.close()
        del d_all
    barrier()

    # ---- parity gate ----------------------------------------------------------------------------------------------
    # (a) the timed occupancy step: sums of the per-rank counters == the oracle's counts for the batch (V from the
    #     oracle's key maths: 1 + |dx| + |dy| + |dz| per ray minus the excluded end voxel; every ray is accepted and hits)
    # (b) a small seeded batch, two steps, NDT: union of the per-rank maps vs the CPU mapper (reference when built)
    gate_n = 3072
    gm_map = make_map("ndt", 0.1, 1 << 30, local_rank)
    connect(gm_map, gate_n)
    for step in range(2):
        lo, hi = 2 * gate_n * step, 2 * gate_n * (step + 1)
        gm_map.exchange_send(sweeps[rank][0][lo:hi])
        gm_map.exchange_integrate()
    gm_map.sync_voxels()
    dump = gm_map.dump()
    gst = gm_map.stats()
    gm_map.close()
    dumps = [None] * world
    dist.gather_object({"dump": dump, "stats": gst}, dumps if rank == 0 else None, dst=0, group=gloo)
    gate = None
    if rank == 0:
        oracle = po.OracleMap(0.1)
        expect_visits = sum(oracle.count_walk_visits(sweeps[r][0], 2) for r in range(world))
        oracle.close()
        gate = {"occupancy_step": {"sum_visits": visits, "oracle_visits": expect_visits, "sum_samples": samples,
                                   "sum_accepted": accepted, "rays": n_step,
                                   "ok": visits == expect_visits and samples == n_step and accepted == n_step}}
        ctor, kind = cpu_mapper("ndt")
        cpu = ctor(0.1)
        for step in range(2):
            for r in range(world):
                cpu.integrate_rays(sweeps[r][0][2 * gate_n * step:2 * gate_n * (step + 1)])
        ref = cpu.dump()
        check = compare_union([d["dump"] for d in dumps], ref, world)
        cpu.close()
        gate["ndt_union_vs_cpu_mapper"] = dict({"cpu_mapper": kind, "rays": 2 * gate_n * world}, **check)
        ndt_total = [sum(c[i] for c in ndt_counts) for i in range(4)]
        gate["config5_counters_vs_single_gpu"] = {
            "sum_visits": ndt_total[1], "single_gpu_visits": one_st["voxel_visits"], "sum_samples": ndt_total[2],
            "single_gpu_samples": one_st["sample_updates"],
            "ok": ndt_total[1] == one_st["voxel_visits"] and ndt_total[2] == one_st["sample_updates"]}
        gate["ok"] = all(v["ok"] for v in gate.values())
        if not gate["ok"]:
            print("PARITY GATE FAILED: " + json.dumps(gate), file=sys.stderr)

    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    alg_bytes = algorithmic_bytes(2, visits, 0, n_step)
    kmax = {k: max(t.get(k, 0.0) for t in all_ktimes) for k in set().union(*all_ktimes)}
    dom_name = max(kmax, key=kmax.get) if kmax else "walkRegions"
    if dom_name == "exWait":  # the wait for the slowest peer is not a kernel of ours to rate
        dom_name = max((k for k in kmax if k != "exWait"), key=kmax.get)
    dom_ms = kmax[dom_name]
    heaviest = max(c[6] for c in all_counts)  # the visits the busiest OWNER applied
    achieved = (8 * heaviest + 44 * n_step / world) / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_how = ncu_traffic(dom_name)
    ndt_value = ndt_rays / (ndt_ms * 1e-3) / 1e6
    single_value = ndt_rays / (single_ms * 1e-3) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": cfg["dtype"], "data": "synthetic",
        "config": {
            "workload": f"{cfg['workload']} per GPU: {world} consecutive sweeps of the moving sensor per step into one "
                        "region-sharded map (fresh each step)",
            "rays_per_step": n_step, "voxel_visits_per_step": visits, "sample_updates_per_step": samples,
            "regions": regions, "resolution_m": cfg["resolution"],
            "parallelism": f"regions sharded over {world} GPUs (owner = (rx + 2 ry + 4 rz) mod {world}); rank r brings sweep r and "
                           "filters / cuts only that sweep; every 32-byte region segment and 96-byte sample record is stored "
                           "into the owner's inbox over NVLink peer memory (CUDA IPC), the 64-byte walk constants follow by "
                           "copy engine; mailbox flags + a bounded device-side wait order the step (ohmb200_exchange_*); no "
                           "NCCL on the data path (torch.distributed carries the IPC handles, the barriers and the timing "
                           "reductions)",
            "segments_per_step": segments_total,
            "visits_share_of_heaviest_owner": heaviest / (visits / world) if visits else None,
            "l2": "map cleared + 512 MiB L2 flush between timed steps (outside the timed spans)",
            "timing": "CUDA events on the launch stream per step around send + integrate, max over ranks, summed over steps; the "
                      "steps are queued back to back with a device-side barrier between the ranks before every span "
                      "(ohmb200_exchange_barrier): no host wait inside or between the spans",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_mine.numel() * 8),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
                "what": "per rank and step: ohmb200_clear + ohmb200_exchange_send(pinned host rays of the OWN sweep) + "
                        "ohmb200_exchange_integrate + counters read + asynchronous download of the occupancy chunks of the "
                        "regions this rank owns; wall clock, max over ranks; bytes are per rank"},
        "gpu_launches": int(launches),
        "kernels_ms_per_step": kmax,
        "kernels_ms_per_step_note": "max over ranks of each kernel's per-step time (profiled pass); exWait = waiting for the "
                                    "slowest peer's records",
        "roofline": {
            "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_how,
            "algorithmic_bytes_per_launch": int(8 * heaviest + 44 * n_step / world), "kernel_ms": dom_ms,
            "peak_source": peak_src, "model": cfg["model"] + ", the heaviest owner's visits",
        },
        "exchange": {
            "nvlink_bytes_per_rank_per_step": int(nvlink_max),
            "what": "bytes the busiest rank writes to its peers in one step: 32 B x segment records + 16 B x sample records whose "
                    "owner is another GPU (peer stores from the cut kernels) + 64 B walk constants x own rays x (N - 1) peers "
                    "(copy engines, beside the cut)",
            "link_floor_ms": nvlink_max / 770e9 * 1e3,
            "link_peak": "770 GB/s per direction per GPU (measured peer copy on this pool, B200_PROFILING.md; 900 nominal)",
            "share_of_step_at_link_peak": (nvlink_max / 770e9 * 1e3) / ms_per_step,
        },
        "config5": {
            "workload": f"BASELINE configs[4] shape: GpuNdtMap (NdtMode::kOccupancy), 0.1 m voxels, moving sensor, regions sharded "
                        f"over {world} GPUs through the routed exchange: {world * (ndt_warm + ndt_steps)} sweeps of the trajectory "
                        f"({world * ndt_warm} warm-up + {world * ndt_steps} timed; the 1000-sweep run is this step repeated)",
            "value": ndt_value, "unit": UNIT, "ms_per_step": ndt_ms / ndt_steps, "steps": ndt_steps, "sweeps_per_step": world,
            "single_gpu": {"value": single_value, "what": "the same sweeps, same order, on ONE of these GPUs (plain path)"},
            "speedup_over_single_gpu": ndt_value / single_value, "efficiency": ndt_value / single_value / world,
        },
        "parity_gate": gate,
        "all_ranks_counters": {"rays_accepted": accepted},
        "clocks": clocks,
    }
    return line


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    # stdout carries the one JSON line and nothing else: whatever a library prints there (NCCL's version banner, for
    # one) is sent to stderr from here on; the line itself is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    line = None
    if world == 1:
        line = run_single(args, torch)
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")  # python objects (IPC handles, counters, the gate's dumps)
        line = run_sharded(args, torch, dist, rank, local_rank, world, gloo)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and line is not None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ohmb200", choices=["ohmb200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="BASELINE.json configs[N-1]: 2 occupancy sweep (default, the driver's line), 3 NDT, 4 TSDF")
    ap.add_argument("--batch", type=int, default=0, help="feed each sweep in batches of this many rays (0 = one call)")
    ap.add_argument("--device-gib", type=float, default=0.0, help="device bytes for the region slabs (0 = per config)")
    ap.add_argument("--cpu-reps", type=int, default=8, help="CPU mapper passes / sweeps for cpu_baseline (0 = skip)")
    ap.add_argument("--pipeline", default="plain", choices=["plain", "exchange"],
                    help="N = 1 A/B: 'exchange' runs the routed-exchange pipeline with world = 1")
    ap.add_argument("--ndt-steps", type=int, default=8, help="N > 1: timed steps of the config-5 (NDT) arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
