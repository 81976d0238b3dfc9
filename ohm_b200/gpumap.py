"""Host-side mirror of ohm's GPU mapper interface over the C ABI (include/ohmb200.h).

``GpuMap`` / ``GpuNdtMap`` / ``GpuTsdfMap`` follow ohmgpu/GpuMap.h:143-303, GpuNdtMap.h:63-110 and
GpuTsdfMap.h:37-80 (same method names in snake_case plus the reference's camelCase spellings), so the parity
tests read like tests/ohmtestgpu/GpuMapTest.cpp.  The C++ facade with the exact reference signatures is
include/ohmb200/GpuMap.hpp; this module exists to drive the library from pytest and bench.py.

All compute happens in libohmb200.so on the GPU; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Params, Stats, KernelTime, ExchangeHandle

LAYER_OCCUPANCY, LAYER_MEAN, LAYER_TRAVERSAL, LAYER_TOUCH_TIME, LAYER_INCIDENT = 0, 1, 2, 3, 4
LAYER_COVARIANCE, LAYER_INTENSITY, LAYER_HIT_MISS, LAYER_TSDF, LAYER_SECONDARY = 5, 6, 7, 8, 9
LAYER_NAMES = ["occupancy", "mean", "traversal", "touch_time", "incident_normal", "covariance", "intensity",
               "hit_miss_count", "tsdf", "secondary_samples"]
LAYER_DTYPES = {
    LAYER_OCCUPANCY: (np.float32, 1),
    LAYER_MEAN: (np.uint32, 2),
    LAYER_TRAVERSAL: (np.float32, 1),
    LAYER_TOUCH_TIME: (np.uint32, 1),
    LAYER_INCIDENT: (np.uint32, 1),
    LAYER_COVARIANCE: (np.float32, 6),
    LAYER_INTENSITY: (np.float32, 2),
    LAYER_HIT_MISS: (np.uint32, 2),
    LAYER_TSDF: (np.float32, 2),
    LAYER_SECONDARY: (np.uint32, 2),  # {f32 m2 | u16 range_mean, u16 count}: raw words
}
MODE_OCCUPANCY, MODE_NDT, MODE_NDT_TM, MODE_TSDF = 0, 1, 2, 3
MODES = {"occupancy": MODE_OCCUPANCY, "ndt": MODE_NDT, "ndt_tm": MODE_NDT_TM, "tsdf": MODE_TSDF}
FILTER_NONE, FILTER_GOOD_RAY, FILTER_CLIP_RANGE, FILTER_CLIP_BOX = 0, 1, 2, 3

# ohm/RayFlag.h:16-60
RF_DEFAULT = 0
RF_END_POINT_AS_FREE = 1 << 0
RF_STOP_ON_FIRST_OCCUPIED = 1 << 1
RF_EXCLUDE_ORIGIN = 1 << 2
RF_EXCLUDE_SAMPLE = 1 << 3
RF_EXCLUDE_RAY = 1 << 4
RF_EXCLUDE_UNOBSERVED = 1 << 5
RF_EXCLUDE_FREE = 1 << 6
RF_EXCLUDE_OCCUPIED = 1 << 7
RF_REVERSE_WALK = 1 << 8


class OhmB200Error(RuntimeError):
    pass


def device_count():
    return _lib.load().ohmb200_device_count()


def default_params(resolution, **overrides):
    p = Params()
    _lib.load().ohmb200_default_params(C.byref(p), float(resolution))
    apply_overrides(p, overrides)
    return p


def apply_overrides(p, overrides):
    for k, v in overrides.items():
        if k in ("region_dim", "origin", "clip_box"):
            arr = getattr(p, k)
            for i in range(len(v)):
                arr[i] = v[i]
        elif k == "layers" and not isinstance(v, int):
            bits = 0
            for layer in v:
                bits |= 1 << layer
            p.layers = bits
        else:
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class GpuMap:
    """ohm::GpuMap over a device-resident occupancy map (ohmgpu/GpuMap.h:143)."""

    mode = "occupancy"

    def __init__(self, resolution=0.1, device_bytes=0, device=0, mode=None, **overrides):
        self.L = _lib.load()
        self.mode = mode or self.mode
        self.params = default_params(resolution, **overrides)
        self.h = self.L.ohmb200_create(C.byref(self.params), MODES[self.mode], int(device_bytes), int(device))
        if not self.h:
            raise OhmB200Error(_lib.last_error())
        self._check(self.L.ohmb200_get_params(self.h, C.byref(self.params)))

    # -- lifetime ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.ohmb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc < 0:
            raise OhmB200Error(f"[{rc}] {_lib.last_error()}")
        return rc

    # -- RayMapper -----------------------------------------------------------------------------------------
    def gpu_ok(self):
        return bool(self.h)

    valid = gpu_ok

    def integrate_rays(self, rays, intensities=None, timestamps=None, ray_flags=RF_DEFAULT):
        """RayMapper::integrateRays: ``rays`` is (2n,3) float64 [origin, sample]* in host memory."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        if intensities is not None:
            intensities = np.ascontiguousarray(intensities, dtype=np.float32)
        if timestamps is not None:
            timestamps = np.ascontiguousarray(timestamps, dtype=np.float64)
        n = self.L.ohmb200_integrate(self.h, _ptr(rays), rays.shape[0], _ptr(intensities), _ptr(timestamps),
                                     int(ray_flags))
        if n == 0 and rays.shape[0] >= 2:
            raise OhmB200Error(_lib.last_error())
        return n

    def integrate_rays_ptr(self, host_ptr, element_count, intensities_ptr=None, timestamps_ptr=None,
                           ray_flags=RF_DEFAULT):
        """Same call with raw host addresses (e.g. pinned buffers)."""
        return self.L.ohmb200_integrate(self.h, host_ptr, element_count, intensities_ptr, timestamps_ptr,
                                        int(ray_flags))

    def integrate_rays_device(self, d_rays_ptr, element_count, d_intensities_ptr=None, d_timestamps_ptr=None,
                              ray_flags=RF_DEFAULT):
        """Rays already resident in HBM (device pointers as ints)."""
        n = self.L.ohmb200_integrate_device(self.h, d_rays_ptr, element_count, d_intensities_ptr, d_timestamps_ptr,
                                            int(ray_flags))
        if n == 0 and element_count >= 2:
            raise OhmB200Error(_lib.last_error())
        return n

    def integrate_secondary(self, rays):
        """ohm::RayMapperSecondarySample::integrateRays on this map (needs LAYER_SECONDARY): ``rays`` = [primary sample,
        secondary sample]* ; returns the number of elements consumed."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] - (rays.shape[0] & 1)
        if n == 0:
            return 0
        done = self.L.ohmb200_integrate_secondary(self.h, _ptr(rays), n)
        if done == 0:
            raise OhmB200Error(_lib.last_error())
        return done

    def sync_voxels(self):
        self._check(self.L.ohmb200_sync(self.h))

    integrateRays = integrate_rays
    syncVoxels = sync_voxels
    gpuOk = gpu_ok

    # -- parameters ----------------------------------------------------------------------------------------
    def set_params(self, **overrides):
        apply_overrides(self.params, overrides)
        self._check(self.L.ohmb200_set_params(self.h, C.byref(self.params)))

    def hit_value(self):
        return self.params.hit_value

    def miss_value(self):
        return self.params.miss_value

    def set_hit_value(self, v):
        self.set_params(hit_value=v)

    def set_miss_value(self, v):
        self.set_params(miss_value=v)

    def first_ray_time(self):
        return self.L.ohmb200_first_ray_time(self.h)

    def set_first_ray_time(self, t):
        """OccupancyMap::setFirstRayTime: the time base of the touch-time layer (restored when a map is loaded)."""
        self._check(self.L.ohmb200_set_first_ray_time(self.h, float(t)))

    # -- map access ----------------------------------------------------------------------------------------
    def region_count(self):
        return self.L.ohmb200_region_count(self.h)

    def region_keys(self):
        n = self.L.ohmb200_enumerate_regions(self.h, None, 0)
        keys = np.zeros((max(n, 1), 3), dtype=np.int16)
        n2 = self.L.ohmb200_enumerate_regions(self.h, keys.ctypes.data_as(C.POINTER(C.c_int16)), n)
        return keys[:min(n, n2)]

    def layers(self):
        return [l for l in range(10) if self.params.layers & (1 << l)]

    def region_layer(self, key, layer):
        return self.region_layers(np.asarray([key], dtype=np.int16), layer)[0]

    def region_layers(self, keys, layer):
        """(count, voxels[, width]) array of one layer for the given region keys: one gather + one D2H."""
        keys = np.ascontiguousarray(keys, dtype=np.int16).reshape(-1, 3)
        dtype, width = LAYER_DTYPES[layer]
        chunk = self.L.ohmb200_region_layer_bytes(self.h, layer)
        if chunk == 0:
            raise OhmB200Error(f"layer {LAYER_NAMES[layer]} is not present")
        out = np.empty(keys.shape[0] * chunk, dtype=np.uint8)
        self._check(self.L.ohmb200_read_regions(self.h, layer, keys.ctypes.data_as(C.POINTER(C.c_int16)),
                                                keys.shape[0], _ptr(out), out.size))
        arr = out.view(dtype)
        nvox = chunk // (np.dtype(dtype).itemsize * width)
        return arr.reshape(keys.shape[0], nvox, width) if width > 1 else arr.reshape(keys.shape[0], nvox)

    def region_layers_async(self, keys, layer, out_ptr, out_bytes):
        """Queue a snapshot + download of one layer of `keys` into host memory at `out_ptr` (pinned for a real overlap);
        returns at once.  Batches integrated afterwards run beside the copy and do not show in it.  The memory must
        stay valid until download_wait()."""
        keys = np.ascontiguousarray(keys, dtype=np.int16).reshape(-1, 3)
        self._check(self.L.ohmb200_read_regions_async(self.h, layer, keys.ctypes.data_as(C.POINTER(C.c_int16)),
                                                      keys.shape[0], C.c_void_p(out_ptr), out_bytes))

    def download_wait(self):
        """Wait for every queued region_layers_async download (GpuMap::syncVoxels' wait on the cache events)."""
        self._check(self.L.ohmb200_download_wait(self.h))

    def rays_query(self, rays, volume_coefficient=1.0):
        """ohm::RaysQuery on the resident map (ohm/RaysQuery.h:42-139): per ray [origin, end], the range to the first
        occupied voxel, the unobserved volume, the terminal OccupancyType and the terminal voxel key.  Returns
        (ranges f64[n], unobserved_volumes f64[n], terminal_states i32[n], terminal_keys i32[n, 6])."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] // 2
        ranges, volumes = np.zeros(n), np.zeros(n)
        states = np.zeros(n, dtype=np.int32)
        keys = np.zeros((n, 6), dtype=np.int32)
        self._check(self.L.ohmb200_rays_query(self.h, _ptr(rays), rays.shape[0], float(volume_coefficient), _ptr(ranges),
                                              _ptr(volumes), _ptr(states), _ptr(keys)))
        return ranges, volumes, states, keys

    def line_keys_query(self, rays):
        """ohm::LineKeysQuery (ohm/LineKeysQuery.h:20-90): the voxel keys along each line.  Returns (result_indices
        u64[n], result_counts u64[n], keys i32[total, 6]); line i owns keys[result_indices[i]:][:result_counts[i]]."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] // 2
        indices, counts = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        total = C.c_size_t(0)
        self._check(self.L.ohmb200_line_keys_query(self.h, _ptr(rays), rays.shape[0], _ptr(indices), _ptr(counts), None,
                                                   0, C.byref(total)))
        keys = np.zeros((total.value, 6), dtype=np.int32)
        if total.value:
            self._check(self.L.ohmb200_line_keys_query(self.h, _ptr(rays), rays.shape[0], _ptr(indices), _ptr(counts),
                                                       _ptr(keys), total.value, C.byref(total)))
        return indices, counts, keys

    def write_region(self, key, layer, data):
        key = np.ascontiguousarray(key, dtype=np.int16)
        data = np.ascontiguousarray(data)
        self._check(self.L.ohmb200_write_region(self.h, key.ctypes.data_as(C.POINTER(C.c_int16)), layer, _ptr(data),
                                                data.nbytes))

    def dump(self):
        """{(rx,ry,rz): {layer: ndarray}} for every resident region (after sync)."""
        self.sync_voxels()
        keys = self.region_keys()
        out = {tuple(int(k) for k in key): {} for key in keys}
        if len(keys) == 0:
            return out
        for layer in self.layers():
            data = self.region_layers(keys, layer)
            for i, key in enumerate(keys):
                out[tuple(int(k) for k in key)][layer] = data[i]
        return out

    def clear(self):
        self._check(self.L.ohmb200_clear(self.h))

    def stats(self):
        s = Stats()
        self._check(self.L.ohmb200_get_stats(self.h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    # -- paging (GpuLayerCache) -----------------------------------------------------------------------------
    def set_region_reserve(self, free_slots):
        """Free region slots guaranteed before every batch (include/ohmb200.h: ohmb200_set_region_reserve)."""
        self._check(self.L.ohmb200_set_region_reserve(self.h, int(free_slots)))

    def remove_region(self, key):
        """MapRegionCache::remove: drop one region (resident or stored) with all its layers."""
        key = np.ascontiguousarray(key, dtype=np.int16)
        self._check(self.L.ohmb200_remove_region(self.h, key.ctypes.data_as(C.POINTER(C.c_int16))))

    def paging_stats(self):
        """{resident, stored, evicted, paged_in}: regions in device memory / in the host store, totals so far."""
        v = [C.c_uint64() for _ in range(4)]
        self._check(self.L.ohmb200_paging_stats(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("resident", "stored", "evicted", "paged_in"), (int(x.value) for x in v)))

    # -- multi-GPU sharding ----------------------------------------------------------------------------------
    def set_partition(self, rank, world):
        """Keep only the regions owned by `rank` of `world` GPUs (include/ohmb200.h: ohmb200_set_partition)."""
        self._check(self.L.ohmb200_set_partition(self.h, int(rank), int(world)))

    def region_owner(self, key, world):
        key = np.ascontiguousarray(key, dtype=np.int16)
        return self.L.ohmb200_region_owner(key.ctypes.data_as(C.POINTER(C.c_int16)), int(world))

    # -- multi-GPU routed exchange (include/ohmb200.h: ohmb200_exchange_*) ---------------------------------------
    def exchange_open(self, rank, world, max_rays_per_rank):
        """Allocate this map's inboxes; returns the 128-byte handle the peers need (bytes)."""
        h = ExchangeHandle()
        self._check(self.L.ohmb200_exchange_open(self.h, int(rank), int(world), int(max_rays_per_rank), C.byref(h)))
        return bytes(h.bytes)

    def exchange_connect(self, handles):
        """`handles`: every rank's handle (bytes), in rank order."""
        arr = (ExchangeHandle * len(handles))()
        for i, raw in enumerate(handles):
            C.memmove(C.byref(arr[i]), bytes(raw), 128)
        self._check(self.L.ohmb200_exchange_connect(self.h, arr, len(handles)))

    def exchange_send(self, rays, intensities=None, timestamps=None, ray_flags=RF_DEFAULT):
        """Phase 1 of a step: this rank's own rays (host arrays) are filtered, cut and routed to the region owners."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        if intensities is not None:
            intensities = np.ascontiguousarray(intensities, dtype=np.float32)
        if timestamps is not None:
            timestamps = np.ascontiguousarray(timestamps, dtype=np.float64)
        n = self.L.ohmb200_exchange_send(self.h, _ptr(rays), rays.shape[0], _ptr(intensities), _ptr(timestamps),
                                         int(ray_flags))
        if n == 0 and rays.shape[0] >= 2:
            raise OhmB200Error(_lib.last_error())
        return n  # (an empty send returns 0 as well: a failure there surfaces in exchange_integrate, "no step pending")

    def exchange_send_ptr(self, host_ptr, element_count, intensities_ptr=None, timestamps_ptr=None, ray_flags=RF_DEFAULT):
        return self.L.ohmb200_exchange_send(self.h, host_ptr, element_count, intensities_ptr, timestamps_ptr, int(ray_flags))

    def exchange_send_device(self, d_rays_ptr, element_count, d_intensities_ptr=None, d_timestamps_ptr=None,
                             ray_flags=RF_DEFAULT):
        n = self.L.ohmb200_exchange_send_device(self.h, d_rays_ptr, element_count, d_intensities_ptr, d_timestamps_ptr,
                                                int(ray_flags))
        if n == 0 and element_count >= 2:
            raise OhmB200Error(_lib.last_error())
        return n

    def exchange_integrate(self):
        """Phase 2: wait (on the device) for every rank's records of the step, integrate what was routed here."""
        self._check(self.L.ohmb200_exchange_integrate(self.h))

    def exchange_barrier(self):
        """Device-side barrier between the ranks (queued; no host wait)."""
        self._check(self.L.ohmb200_exchange_barrier(self.h))

    def exchange_last_counts(self, world):
        """(segment records, sample records) this rank sent to each owner in the last step."""
        seg = (C.c_uint32 * world)()
        smp = (C.c_uint32 * world)()
        self._check(self.L.ohmb200_exchange_last_counts(self.h, seg, smp, world))
        return list(seg), list(smp)

    def exchange_close(self):
        self._check(self.L.ohmb200_exchange_close(self.h))

    # -- measurement ---------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.L.ohmb200_set_stream(self.h, cuda_stream))

    def set_profiling(self, enabled):
        self._check(self.L.ohmb200_set_profiling(self.h, int(bool(enabled))))

    def kernel_times(self, reset=True):
        arr = (KernelTime * 40)()
        n = self._check(self.L.ohmb200_kernel_times(self.h, arr, 40, int(reset)))
        return {arr[i].name.decode(): {"ms": arr[i].ms, "launches": int(arr[i].launches)} for i in range(n)}


class GpuNdtMap(GpuMap):
    """ohm::GpuNdtMap (ohmgpu/GpuNdtMap.h:63): occupancy + voxel mean + covariance."""

    mode = "ndt"

    def __init__(self, resolution=0.1, traversability=False, **kw):
        super().__init__(resolution, mode="ndt_tm" if traversability else "ndt", **kw)

    def set_sensor_noise(self, noise):
        self.set_params(sensor_noise=noise)

    def sensor_noise(self):
        return self.params.sensor_noise


class GpuTsdfMap(GpuMap):
    """ohm::GpuTsdfMap (ohmgpu/GpuTsdfMap.h:37)."""

    mode = "tsdf"

    def set_tsdf_options(self, max_weight=None, default_truncation_distance=None, dropoff_epsilon=None,
                         sparsity_compensation_factor=None):
        kw = {}
        if max_weight is not None:
            kw["tsdf_max_weight"] = max_weight
        if default_truncation_distance is not None:
            kw["tsdf_trunc"] = default_truncation_distance
        if dropoff_epsilon is not None:
            kw["tsdf_dropoff"] = dropoff_epsilon
        if sparsity_compensation_factor is not None:
            kw["tsdf_sparsity"] = sparsity_compensation_factor
        self.set_params(**kw)


def open_exchange(maps, max_rays_per_rank):
    """Connect `maps` (one per rank, in rank order, all in this process) into one routed exchange."""
    handles = [m.exchange_open(r, len(maps), max_rays_per_rank) for r, m in enumerate(maps)]
    for m in maps:
        m.exchange_connect(handles)
    return handles


def exchange_step(maps, batches, ray_flags=RF_DEFAULT):
    """One step of the exchange driven from a single process: every rank sends, then every rank integrates.
    `batches[r]` = (rays, intensities, timestamps) of rank r (rays may be empty)."""
    for m, (rays, intensities, timestamps) in zip(maps, batches):
        m.exchange_send(rays, intensities, timestamps, ray_flags)
    for m in maps:
        m.exchange_integrate()
