"""ctypes binding of oracle/_ref/libohm_ref.so — the REFERENCE's own CPU mappers (ohm::RayMapperOccupancy / Ndt /
Tsdf on an ohm::OccupancyMap), compiled unmodified from /root/reference by `make -C oracle ref`.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline / --impl reference).  The .so is built only where
/root/reference exists; it is git-ignored but travels to the GPU box with the snapshot.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import pyoracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libohm_ref.so")
REFERENCE = "/root/reference"
MODES = {"occupancy": 0, "ndt": 1, "ndt_tm": 2, "tsdf": 3}


def available(build=True):
    """True when the reference library exists (building it first if the reference sources are present)."""
    if os.path.exists(LIB_PATH):
        return True
    if build and os.path.isdir(os.path.join(REFERENCE, "ohm")):
        try:
            subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "ref"])
        except Exception:
            return False
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise ImportError("oracle/_ref/libohm_ref.so is not built (needs /root/reference)")
    L = C.CDLL(LIB_PATH)
    vp, dp, fp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float)
    L.ref_map_create.argtypes = [C.POINTER(po.Params), C.c_int]
    L.ref_map_create.restype = vp
    L.ref_map_destroy.argtypes = [vp]
    L.ref_map_set_params.argtypes = [vp, C.POINTER(po.Params)]
    L.ref_mapper_valid.argtypes = [vp]
    L.ref_integrate.argtypes = [vp, dp, C.c_size_t, fp, dp, C.c_uint]
    L.ref_integrate.restype = C.c_size_t
    L.ref_first_ray_time.argtypes = [vp]
    L.ref_first_ray_time.restype = C.c_double
    L.ref_region_count.argtypes = [vp]
    L.ref_region_count.restype = C.c_size_t
    L.ref_region_keys.argtypes = [vp, C.POINTER(C.c_int16), C.c_size_t]
    L.ref_region_keys.restype = C.c_size_t
    L.ref_region_layer.argtypes = [vp, C.POINTER(C.c_int16), C.c_int, vp, C.c_size_t]
    L.ref_region_layer.restype = C.c_size_t
    L.ref_walk_segment.argtypes = [vp, dp, dp, C.c_uint, C.POINTER(C.c_int32), dp, dp, C.c_size_t]
    L.ref_walk_segment.restype = C.c_size_t
    L.ref_rays_query.argtypes = [vp, dp, C.c_size_t, C.c_double, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int32)]
    L.ref_rays_query.restype = C.c_size_t
    L.ref_line_keys_query.argtypes = [vp, dp, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_int32), C.c_size_t]
    L.ref_line_keys_query.restype = C.c_size_t
    L.ref_integrate_secondary.argtypes = [vp, dp, C.c_size_t]
    L.ref_integrate_secondary.restype = C.c_size_t
    L.ref_save.argtypes = [vp, C.c_char_p]
    L.ref_save.restype = C.c_int
    L.ref_load.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
    L.ref_load.restype = vp
    L.ref_map_header.argtypes = [vp, dp, dp, C.POINTER(C.c_int), dp, dp, dp, dp, C.POINTER(C.c_uint)]
    _lib = L
    return L


class ReferenceMap:
    """Same surface as pyoracle.OracleMap, backed by the real ohm classes."""

    def __init__(self, resolution=0.1, mode="occupancy", **overrides):
        self.L = lib()
        self.params = po.default_params(resolution, **overrides)
        if mode in ("ndt", "ndt_tm"):
            self.params.layers |= (1 << po.LAYER_MEAN) | (1 << po.LAYER_COVARIANCE)
        if mode == "ndt_tm":
            self.params.layers |= (1 << po.LAYER_INTENSITY) | (1 << po.LAYER_HIT_MISS)
        if mode == "tsdf":
            self.params.layers = 1 << po.LAYER_TSDF
        self.mode = mode
        self.h = self.L.ref_map_create(C.byref(self.params), MODES[mode])
        assert self.L.ref_mapper_valid(self.h)

    def close(self):
        if self.h:
            self.L.ref_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **overrides):
        po.apply_overrides(self.params, overrides)
        self.L.ref_map_set_params(self.h, C.byref(self.params))

    def integrate_rays(self, rays, intensities=None, timestamps=None, ray_flags=0):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] - (rays.shape[0] & 1)
        ip = tp = None
        if intensities is not None:
            intensities = np.ascontiguousarray(intensities, dtype=np.float32)
            ip = intensities.ctypes.data_as(C.POINTER(C.c_float))
        if timestamps is not None:
            timestamps = np.ascontiguousarray(timestamps, dtype=np.float64)
            tp = timestamps.ctypes.data_as(C.POINTER(C.c_double))
        return self.L.ref_integrate(self.h, rays.ctypes.data_as(C.POINTER(C.c_double)), n, ip, tp, int(ray_flags))

    def first_ray_time(self):
        return self.L.ref_first_ray_time(self.h)

    def layers(self):
        return [l for l in range(10) if self.params.layers & (1 << l)]

    def region_keys(self):
        n = self.L.ref_region_count(self.h)
        keys = np.zeros((max(n, 1), 3), dtype=np.int16)
        self.L.ref_region_keys(self.h, keys.ctypes.data_as(C.POINTER(C.c_int16)), n)
        return keys[:n]

    def region_layer(self, key, layer):
        key = np.ascontiguousarray(key, dtype=np.int16)
        dtype, width = po.LAYER_DTYPES[layer]
        d = self.params.region_dim
        nvox = d[0] * d[1] * d[2]
        out = np.empty(nvox * width, dtype=dtype)
        got = self.L.ref_region_layer(self.h, key.ctypes.data_as(C.POINTER(C.c_int16)), layer, C.c_void_p(out.ctypes.data),
                                      out.nbytes)
        if got == 0:
            return None
        assert got == out.nbytes, (got, out.nbytes)
        return out.reshape(nvox, width) if width > 1 else out

    def dump(self):
        out = {}
        for key in self.region_keys():
            out[tuple(int(k) for k in key)] = {l: self.region_layer(key, l) for l in self.layers()}
        return out

    def integrate_secondary(self, rays):
        """ohm::RayMapperSecondarySample::integrateRays on this map."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] - (rays.shape[0] & 1)
        return self.L.ref_integrate_secondary(self.h, rays.ctypes.data_as(C.POINTER(C.c_double)), n)

    def save(self, path):
        """ohm::save(path, map) (ohm/MapSerialise.cpp:595-648)."""
        rc = self.L.ref_save(self.h, str(path).encode())
        assert rc == 0, f"ohm::save failed: {rc}"

    @classmethod
    def load(cls, path, layers):
        """ohm::load(path, map) into a fresh reference map; `layers` = the layer ids to expose through dump()."""
        err = C.c_int(0)
        h = lib().ref_load(str(path).encode(), C.byref(err))
        if not h:
            raise RuntimeError(f"ohm::load failed: {err.value}")
        self = cls.__new__(cls)
        self.L, self.h, self.mode = lib(), h, "loaded"
        res, thr, hit, miss, frt = (C.c_double() for _ in range(5))
        origin, dim, flags = (C.c_double * 3)(), (C.c_int * 3)(), C.c_uint()
        self.L.ref_map_header(h, C.byref(res), origin, dim, C.byref(frt), C.byref(thr), C.byref(hit), C.byref(miss),
                              C.byref(flags))
        self.params = po.default_params(res.value, origin=tuple(origin), region_dim=tuple(dim), layers=list(layers))
        self.header = dict(resolution=res.value, origin=tuple(origin), region_dim=tuple(dim), first_ray_time=frt.value,
                           threshold_value=thr.value, hit_value=hit.value, miss_value=miss.value, flags=flags.value)
        return self

    def rays_query(self, rays, volume_coefficient=1.0):
        """ohm::RaysQuery on the reference map: (ranges, unobserved_volumes, terminal_states, terminal_keys)."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] // 2
        ranges, volumes = np.zeros(n), np.zeros(n)
        states = np.zeros(n, dtype=np.int32)
        keys = np.zeros((n, 6), dtype=np.int32)
        dp = C.POINTER(C.c_double)
        got = self.L.ref_rays_query(self.h, rays.ctypes.data_as(dp), rays.shape[0], float(volume_coefficient),
                                    ranges.ctypes.data_as(dp), volumes.ctypes.data_as(dp),
                                    states.ctypes.data_as(C.POINTER(C.c_int)), keys.ctypes.data_as(C.POINTER(C.c_int32)))
        assert got == n
        return ranges, volumes, states, keys

    def line_keys_query(self, rays, cap=1 << 22):
        """ohm::LineKeysQuery on the reference map: (result_indices u64[n], result_counts u64[n], keys i32[total, 6])."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 3)
        n = rays.shape[0] // 2
        indices, counts = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
        keys = np.zeros((cap, 6), dtype=np.int32)
        total = self.L.ref_line_keys_query(self.h, rays.ctypes.data_as(C.POINTER(C.c_double)), rays.shape[0],
                                           indices.ctypes.data_as(C.POINTER(C.c_uint64)),
                                           counts.ctypes.data_as(C.POINTER(C.c_uint64)),
                                           keys.ctypes.data_as(C.POINTER(C.c_int32)), cap)
        assert total <= cap
        return indices, counts, keys[:total]

    def walk_segment(self, start, end, walk_flags=0, cap=1 << 16):
        start = np.ascontiguousarray(start, dtype=np.float64)
        end = np.ascontiguousarray(end, dtype=np.float64)
        keys = np.zeros((cap, 6), dtype=np.int32)
        enter = np.zeros(cap)
        exit_ = np.zeros(cap)
        dp = C.POINTER(C.c_double)
        n = self.L.ref_walk_segment(self.h, start.ctypes.data_as(dp), end.ctypes.data_as(dp), walk_flags,
                                    keys.ctypes.data_as(C.POINTER(C.c_int32)), enter.ctypes.data_as(dp),
                                    exit_.ctypes.data_as(dp), cap)
        return keys[:n], enter[:n], exit_[:n]
