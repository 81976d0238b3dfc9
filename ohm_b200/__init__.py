"""ohm_b200 — B200-native (sm_100a) batched ray integration behind ohm's RayMapper/GpuMap interface.

The product is libohmb200.so (CUDA, C ABI in include/ohmb200.h); this package is its host-side driver.
"""
from .gpumap import (  # noqa: F401
    GpuMap, GpuNdtMap, GpuTsdfMap, OhmB200Error, device_count, default_params,
    LAYER_OCCUPANCY, LAYER_MEAN, LAYER_TRAVERSAL, LAYER_TOUCH_TIME, LAYER_INCIDENT, LAYER_COVARIANCE,
    LAYER_INTENSITY, LAYER_HIT_MISS, LAYER_TSDF,
    RF_DEFAULT, RF_END_POINT_AS_FREE, RF_STOP_ON_FIRST_OCCUPIED, RF_EXCLUDE_ORIGIN, RF_EXCLUDE_SAMPLE,
    RF_EXCLUDE_RAY, RF_EXCLUDE_UNOBSERVED, RF_EXCLUDE_FREE, RF_EXCLUDE_OCCUPIED, RF_REVERSE_WALK,
    FILTER_NONE, FILTER_GOOD_RAY, FILTER_CLIP_RANGE, FILTER_CLIP_BOX,
)
