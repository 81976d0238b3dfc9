"""ctypes binding of libohmb200.so (the C ABI in include/ohmb200.h).

The CUDA library is mandatory: importing this module raises if it is missing — there is no CPU fallback and
nothing here ever touches oracle/.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OHMB200_LIB") or os.path.join(HERE, "libohmb200.so")  # OHMB200_LIB: instrumented builds

LAYER_COUNT = 10


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
    _fields_ = [(n, C.c_uint64) for n in (
        "rays_in", "rays_accepted", "voxel_visits", "sample_updates", "ordered_records", "regions",
        "region_capacity", "batches", "kernel_launches", "sample_voxels", "owned_visits")]


class ExchangeHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 128)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_double), ("launches", C.c_uint64)]


# Every symbol include/ohmb200.h declares: (name, restype, argtypes)
_vp = C.c_void_p
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_kp = C.POINTER(C.c_int16)
SYMBOLS = [
    ("ohmb200_device_count", C.c_int, []),
    ("ohmb200_default_params", None, [C.POINTER(Params), C.c_double]),
    ("ohmb200_create", _vp, [C.POINTER(Params), C.c_int, C.c_size_t, C.c_int]),
    ("ohmb200_destroy", None, [_vp]),
    ("ohmb200_set_params", C.c_int, [_vp, C.POINTER(Params)]),
    ("ohmb200_get_params", C.c_int, [_vp, C.POINTER(Params)]),
    ("ohmb200_integrate", C.c_size_t, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_uint]),
    ("ohmb200_integrate_device", C.c_size_t, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_uint]),
    ("ohmb200_sync", C.c_int, [_vp]),
    ("ohmb200_region_count", C.c_size_t, [_vp]),
    ("ohmb200_enumerate_regions", C.c_size_t, [_vp, _kp, C.c_size_t]),
    ("ohmb200_region_layer_bytes", C.c_size_t, [_vp, C.c_int]),
    ("ohmb200_read_region", C.c_int, [_vp, _kp, C.c_int, _vp, C.c_size_t]),
    ("ohmb200_read_regions", C.c_int, [_vp, C.c_int, _kp, C.c_size_t, _vp, C.c_size_t]),
    ("ohmb200_read_regions_async", C.c_int, [_vp, C.c_int, _kp, C.c_size_t, _vp, C.c_size_t]),
    ("ohmb200_download_wait", C.c_int, [_vp]),
    ("ohmb200_integrate_secondary", C.c_size_t, [_vp, _vp, C.c_size_t]),
    ("ohmb200_integrate_secondary_device", C.c_size_t, [_vp, _vp, C.c_size_t]),
    ("ohmb200_rays_query", C.c_int, [_vp, _vp, C.c_size_t, C.c_double, _vp, _vp, _vp, _vp]),
    ("ohmb200_rays_query_device", C.c_int, [_vp, _vp, C.c_size_t, C.c_double, _vp, _vp, _vp, _vp]),
    ("ohmb200_line_keys_query", C.c_int, [_vp, _vp, C.c_size_t, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("ohmb200_write_region", C.c_int, [_vp, _kp, C.c_int, _vp, C.c_size_t]),
    ("ohmb200_clear", C.c_int, [_vp]),
    ("ohmb200_first_ray_time", C.c_double, [_vp]),
    ("ohmb200_set_first_ray_time", C.c_int, [_vp, C.c_double]),
    ("ohmb200_get_stats", C.c_int, [_vp, C.POINTER(Stats)]),
    ("ohmb200_set_stream", C.c_int, [_vp, _vp]),
    ("ohmb200_set_partition", C.c_int, [_vp, C.c_int, C.c_int]),
    ("ohmb200_region_owner", C.c_int, [_kp, C.c_int]),
    ("ohmb200_set_profiling", C.c_int, [_vp, C.c_int]),
    ("ohmb200_kernel_times", C.c_int, [_vp, C.POINTER(KernelTime), C.c_int, C.c_int]),
    ("ohmb200_set_region_reserve", C.c_int, [_vp, C.c_uint32]),
    ("ohmb200_remove_region", C.c_int, [_vp, _kp]),
    ("ohmb200_paging_stats", C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64)]),
    ("ohmb200_exchange_open", C.c_int, [_vp, C.c_int, C.c_int, C.c_size_t, C.POINTER(ExchangeHandle)]),
    ("ohmb200_exchange_connect", C.c_int, [_vp, C.POINTER(ExchangeHandle), C.c_int]),
    ("ohmb200_exchange_send", C.c_size_t, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_uint]),
    ("ohmb200_exchange_send_device", C.c_size_t, [_vp, _vp, C.c_size_t, _vp, _vp, C.c_uint]),
    ("ohmb200_exchange_integrate", C.c_int, [_vp]),
    ("ohmb200_exchange_close", C.c_int, [_vp]),
    ("ohmb200_exchange_barrier", C.c_int, [_vp]),
    ("ohmb200_exchange_last_counts", C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int]),
    ("ohmb200_last_error", C.c_char_p, []),
    ("ohmb200_version", C.c_char_p, []),
]

_lib = None


def load():
    """Load libohmb200.so; raises ImportError when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ohm_b200.build` (nvcc, sm_100a). "
            "ohm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error():
    return load().ohmb200_last_error().decode()
