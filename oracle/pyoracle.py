"""ctypes binding of the CPU oracle (oracle/libohm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under ohm_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libohm_oracle.so")

LAYER_OCCUPANCY, LAYER_MEAN, LAYER_TRAVERSAL, LAYER_TOUCH_TIME, LAYER_INCIDENT = 0, 1, 2, 3, 4
LAYER_COVARIANCE, LAYER_INTENSITY, LAYER_HIT_MISS, LAYER_TSDF, LAYER_SECONDARY = 5, 6, 7, 8, 9
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
    LAYER_SECONDARY: (np.uint32, 2),  # {f32 m2 | u16 range_mean, u16 count}, compared as raw words
}
FILTER_NONE, FILTER_GOOD_RAY, FILTER_CLIP_RANGE, FILTER_CLIP_BOX = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [
        ("resolution", C.c_double),
        ("region_dim", C.c_int32 * 3),
        ("origin", C.c_double * 3),
        ("hit_value", C.c_float),
        ("miss_value", C.c_float),
        ("min_value", C.c_float),
        ("max_value", C.c_float),
        ("threshold_value", C.c_float),
        ("saturate_min", C.c_int32),
        ("saturate_max", C.c_int32),
        ("layers", C.c_uint32),
        ("filter_kind", C.c_int32),
        ("filter_range", C.c_double),
        ("clip_box", C.c_double * 6),
        ("sensor_noise", C.c_float),
        ("adaptation_rate", C.c_float),
        ("reinit_threshold", C.c_float),
        ("reinit_count", C.c_uint32),
        ("sample_threshold", C.c_uint32),
        ("initial_intensity_cov", C.c_float),
        ("ndt_tm", C.c_int32),
        ("tsdf_max_weight", C.c_float),
        ("tsdf_trunc", C.c_float),
        ("tsdf_dropoff", C.c_float),
        ("tsdf_sparsity", C.c_float),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays_in", C.c_uint64),
        ("rays_accepted", C.c_uint64),
        ("voxel_visits", C.c_uint64),
        ("sample_updates", C.c_uint64),
    ]


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "ohm_oracle.c")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libohm_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp = C.POINTER(C.c_double)
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int32)
    vp = C.c_void_p
    L.oracle_default_params.argtypes = [C.POINTER(Params), C.c_double]
    L.oracle_map_create.argtypes = [C.POINTER(Params)]
    L.oracle_map_create.restype = vp
    L.oracle_map_destroy.argtypes = [vp]
    L.oracle_map_set_params.argtypes = [vp, C.POINTER(Params)]
    L.oracle_map_stats.argtypes = [vp, C.POINTER(Stats)]
    L.oracle_first_ray_time.argtypes = [vp]
    L.oracle_first_ray_time.restype = C.c_double
    L.oracle_voxel_key.argtypes = [vp, dp, ip]
    L.oracle_voxel_key.restype = C.c_int
    L.oracle_voxel_centre.argtypes = [vp, ip, dp]
    L.oracle_walk_segment.argtypes = [vp, dp, dp, C.c_uint, ip, dp, dp, C.c_size_t]
    L.oracle_walk_segment.restype = C.c_size_t
    for name in ("oracle_integrate_occupancy", "oracle_integrate_ndt", "oracle_integrate_tsdf"):
        f = getattr(L, name)
        f.argtypes = [vp, dp, C.c_size_t, fp, dp, C.c_uint]
        f.restype = C.c_size_t
    L.oracle_rays_query.argtypes = [vp, dp, C.c_size_t, C.c_double, dp, dp, C.POINTER(C.c_int), ip]
    L.oracle_rays_query.restype = C.c_size_t
    L.oracle_integrate_secondary.argtypes = [vp, dp, C.c_size_t]
    L.oracle_integrate_secondary.restype = C.c_size_t
    L.oracle_count_walk_visits.argtypes = [vp, dp, C.c_size_t, C.c_uint]
    L.oracle_count_walk_visits.restype = C.c_uint64
    L.oracle_region_count.argtypes = [vp]
    L.oracle_region_count.restype = C.c_size_t
    L.oracle_region_keys.argtypes = [vp, C.POINTER(C.c_int16), C.c_size_t]
    L.oracle_region_keys.restype = C.c_size_t
    L.oracle_region_layer.argtypes = [vp, C.POINTER(C.c_int16), C.c_int]
    L.oracle_region_layer.restype = vp
    L.oracle_sub_voxel_update.argtypes = [C.c_uint32, C.c_uint32, dp, C.c_double]
    L.oracle_sub_voxel_update.restype = C.c_uint32
    L.oracle_sub_voxel_to_local.argtypes = [C.c_uint32, C.c_double, dp]
    L.oracle_update_incident_normal.argtypes = [C.c_uint32, fp, C.c_uint32]
    L.oracle_update_incident_normal.restype = C.c_uint32
    L.oracle_decode_normal.argtypes = [C.c_uint32, fp]
    L.oracle_encode_normal.argtypes = [fp]
    L.oracle_encode_normal.restype = C.c_uint32
    L.oracle_encode_touch_time.argtypes = [C.c_double, C.c_double]
    L.oracle_encode_touch_time.restype = C.c_uint32
    L.oracle_calculate_hit_with_covariance.argtypes = [
        fp, fp, dp, dp, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32]
    L.oracle_calculate_hit_with_covariance.restype = C.c_int
    L.oracle_calculate_miss_ndt.argtypes = [
        fp, fp, C.POINTER(C.c_int), dp, dp, dp, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint32]
    L.oracle_calculate_tsdf.argtypes = [dp, dp, dp, C.c_float, C.c_float, C.c_float, C.c_float, fp, fp]
    L.oracle_calculate_tsdf.restype = C.c_int
    L.oracle_occupancy_adjust_miss.argtypes = [fp] + [C.c_float] * 6 + [C.c_int]
    L.oracle_occupancy_adjust_hit.argtypes = [fp] + [C.c_float] * 6 + [C.c_int]
    _lib = L
    return L


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_params(resolution, **overrides):
    p = Params()
    lib().oracle_default_params(C.byref(p), float(resolution))
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


class OracleMap:
    """CPU RayMapperOccupancy / RayMapperNdt / RayMapperTsdf over an oracle map (mode picks the mapper)."""

    def __init__(self, resolution=0.1, mode="occupancy", **overrides):
        self.L = lib()
        self.params = default_params(resolution, **overrides)
        self.mode = mode
        self.h = self.L.oracle_map_create(C.byref(self.params))

    def close(self):
        if self.h:
            self.L.oracle_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **overrides):
        apply_overrides(self.params, overrides)
        self.L.oracle_map_set_params(self.h, C.byref(self.params))

    def integrate_rays(self, rays, intensities=None, timestamps=None, ray_flags=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] - (rays.shape[0] & 1)
        ip = tp = None
        if intensities is not None:
            intensities = np.ascontiguousarray(intensities, dtype=np.float32)
            ip = _fptr(intensities)
        if timestamps is not None:
            timestamps = np.ascontiguousarray(timestamps, dtype=np.float64)
            tp = _dptr(timestamps)
        fn = {
            "occupancy": self.L.oracle_integrate_occupancy,
            "ndt": self.L.oracle_integrate_ndt,
            "ndt_tm": self.L.oracle_integrate_ndt,
            "tsdf": self.L.oracle_integrate_tsdf,
        }[self.mode]
        return fn(self.h, _dptr(rays), n, ip, tp, int(ray_flags))

    def rays_query(self, rays, volume_coefficient=1.0):
        """ohm::RaysQuery (ohm/RaysQuery.cpp:109-199): (ranges f64[n], unobserved_volumes f64[n], terminal_states i32[n],
        terminal_keys i32[n, 6])."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] // 2
        ranges, volumes = np.zeros(n), np.zeros(n)
        states = np.zeros(n, dtype=np.int32)
        keys = np.zeros((n, 6), dtype=np.int32)
        got = self.L.oracle_rays_query(self.h, _dptr(rays), rays.shape[0], float(volume_coefficient), _dptr(ranges),
                                       _dptr(volumes), states.ctypes.data_as(C.POINTER(C.c_int)),
                                       keys.ctypes.data_as(C.POINTER(C.c_int32)))
        assert got == n
        return ranges, volumes, states, keys

    def line_keys_query(self, rays):
        """ohm::LineKeysQuery (ohm/LineKeysQuery.cpp:103-123) = calculateSegmentKeys per line, end voxel included:
        (result_indices u64[n], result_counts u64[n], keys i32[total, 6])."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] // 2
        indices, counts = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        chunks = []
        at = 0
        for i in range(n):
            keys, _, _ = self.walk_segment(rays[2 * i], rays[2 * i + 1], 0)
            indices[i], counts[i] = at, len(keys)
            at += len(keys)
            chunks.append(keys.copy())
        return indices, counts, (np.concatenate(chunks) if chunks else np.zeros((0, 6), dtype=np.int32))

    def integrate_secondary(self, rays):
        """RayMapperSecondarySample::integrateRays (ohm/RayMapperSecondarySample.cpp:37-74)."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        return self.L.oracle_integrate_secondary(self.h, _dptr(rays), rays.shape[0] - (rays.shape[0] & 1))

    def count_walk_visits(self, rays, walk_flags=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        return int(self.L.oracle_count_walk_visits(self.h, _dptr(rays), rays.shape[0], int(walk_flags)))

    def stats(self):
        s = Stats()
        self.L.oracle_map_stats(self.h, C.byref(s))
        return {k: int(getattr(s, k)) for k, _ in Stats._fields_}

    def first_ray_time(self):
        return self.L.oracle_first_ray_time(self.h)

    def voxel_key(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        key = np.zeros(6, dtype=np.int32)
        ok = self.L.oracle_voxel_key(self.h, _dptr(p), key.ctypes.data_as(C.POINTER(C.c_int32)))
        return key if ok else None

    def voxel_centre(self, key):
        key = np.ascontiguousarray(key, dtype=np.int32)
        out = np.zeros(3, dtype=np.float64)
        self.L.oracle_voxel_centre(self.h, key.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(out))
        return out

    def walk_segment(self, start, end, walk_flags=0, cap=1 << 16):
        start = np.ascontiguousarray(start, dtype=np.float64)
        end = np.ascontiguousarray(end, dtype=np.float64)
        keys = np.zeros((cap, 6), dtype=np.int32)
        enter = np.zeros(cap, dtype=np.float64)
        exit_ = np.zeros(cap, dtype=np.float64)
        n = self.L.oracle_walk_segment(
            self.h, _dptr(start), _dptr(end), walk_flags, keys.ctypes.data_as(C.POINTER(C.c_int32)),
            _dptr(enter), _dptr(exit_), cap)
        assert n <= cap
        return keys[:n], enter[:n], exit_[:n]

    def region_keys(self):
        n = self.L.oracle_region_count(self.h)
        keys = np.zeros((max(n, 1), 3), dtype=np.int16)
        self.L.oracle_region_keys(self.h, keys.ctypes.data_as(C.POINTER(C.c_int16)), n)
        return keys[:n]

    def region_layer(self, key, layer):
        key = np.ascontiguousarray(key, dtype=np.int16)
        ptr = self.L.oracle_region_layer(self.h, key.ctypes.data_as(C.POINTER(C.c_int16)), layer)
        if not ptr:
            return None
        dtype, width = LAYER_DTYPES[layer]
        d = self.params.region_dim
        nvox = d[0] * d[1] * d[2]
        buf = (C.c_char * (nvox * width * np.dtype(dtype).itemsize)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype).copy()
        return arr.reshape(nvox, width) if width > 1 else arr

    def layers(self):
        return [l for l in range(10) if self.params.layers & (1 << l)]

    def dump(self):
        """{(rx,ry,rz): {layer: ndarray}} for every region."""
        out = {}
        for key in self.region_keys():
            out[tuple(int(k) for k in key)] = {l: self.region_layer(key, l) for l in self.layers()}
        return out
