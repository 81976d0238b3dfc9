"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/ohmb200.h
declares, parameter defaults match ohm's, and the product never reaches into oracle/."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

import ohm_b200
from ohm_b200 import _lib, gpumap as gm
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ohmb200.h")).read()
    return sorted(set(re.findall(r"OHMB200_API[^;]*?\b(ohmb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ohmb200.h but not exported"
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(declared) == bound, f"ctypes table and header disagree: {set(declared) ^ bound}"
    assert b"sm_100a" in lib.ohmb200_version()


def test_only_abi_symbols_are_exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    extra = {s for s in exported if not s.startswith("ohmb200_")}
    assert not extra, f"non-ABI symbols leak from libohmb200.so: {sorted(extra)[:5]}"


def test_default_params_match_ohm_defaults_and_oracle():
    p = gm.default_params(0.1)
    o = po.default_params(0.1)
    assert C.sizeof(p) == C.sizeof(o)
    assert bytes(p) == bytes(o)
    assert list(p.region_dim) == [32, 32, 32] and p.filter_kind == gm.FILTER_GOOD_RAY and p.filter_range == 1e10
    assert abs(p.hit_value - 2.1972246) < 1e-6 and abs(p.miss_value + 0.2006707) < 1e-6
    assert p.sample_threshold == 3 and p.reinit_count == 100 and abs(p.reinit_threshold + 1.3862944) < 1e-6


def test_no_gpu_means_loud_failure_not_fallback():
    if ohm_b200.device_count() > 0:
        return
    try:
        ohm_b200.GpuMap(0.1)
    except ohm_b200.OhmB200Error as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("GpuMap constructed without a GPU")


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ohm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "ohm_oracle" not in text and "import oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower().replace("test oracle", "")
    deps = subprocess.check_output(["ldd", _lib.LIB_PATH]).decode()
    assert "oracle" not in deps


def test_region_owner_partitions_regions():
    lib = _lib.load()
    keys = np.array([[x, y, z] for x in range(-6, 7) for y in range(-5, 6) for z in range(-2, 3)], dtype=np.int16)
    for world in (1, 2, 4, 8):
        owners = np.array([lib.ohmb200_region_owner(k.ctypes.data_as(C.POINTER(C.c_int16)), world) for k in keys])
        assert owners.min() >= 0 and owners.max() < world
        counts = np.bincount(owners, minlength=world)
        assert counts.min() > 0.5 * len(keys) / world, counts   # reasonably balanced
        # the formula in the header; with 8 owners face neighbours never share one
        for k, o in zip(keys[::37], owners[::37]):
            assert o == (int(k[0]) + 2 * int(k[1]) + 4 * int(k[2])) % world
        if world == 8:
            a = lib.ohmb200_region_owner(np.array([2, 4, 0], dtype=np.int16).ctypes.data_as(C.POINTER(C.c_int16)), world)
            for d in ([1, 0, 0], [0, 1, 0], [0, 0, 1]):
                n = np.array([2 + d[0], 4 + d[1], d[2]], dtype=np.int16)
                assert lib.ohmb200_region_owner(n.ctypes.data_as(C.POINTER(C.c_int16)), world) != a


def test_reference_side_binding_defines_the_ohmgpu_classes():
    """Where it was built (needs /root/reference): tests/binding/GpuMapB200.cpp, compiled against the reference's own
    ohmgpu/GpuMap.h, GpuNdtMap.h, GpuTsdfMap.h and ohm/MapRegionCache.h, defines the classes OhmAppGpu links against and
    reaches the device only through the C ABI (no compute here: symbols only)."""
    import pytest
    path = os.path.join(ROOT, "oracle", "_ref", "libohm_b200_binding.so")
    if not os.path.exists(path):
        pytest.skip("the binding is built only where /root/reference exists")
    out = subprocess.check_output(["nm", "-DC", path]).decode()
    defined = [line for line in out.splitlines() if " T " in line or " W " in line]
    for symbol in ("ohm::GpuMap::GpuMap(ohm::OccupancyMap*, bool, unsigned int, unsigned long)", "ohm::GpuMap::integrateRays(",
                   "ohm::GpuMap::syncVoxels()", "ohm::GpuMap::syncVoxels(std::vector<int", "ohm::GpuMap::setRayFilter(",
                   "ohm::GpuMap::gpuCache() const", "ohm::GpuNdtMap::GpuNdtMap(", "ohm::GpuNdtMap::setSensorNoise(float)",
                   "ohm::GpuTsdfMap::GpuTsdfMap(", "ohm::GpuTsdfMap::setTsdfOptions(", "ohm::gpumap::enableGpu(ohm::OccupancyMap&)",
                   "ohm::gpumap::sync(ohm::OccupancyMap&)", "ohm::gpumap::gpuCache(ohm::OccupancyMap&)",
                   "ohm::GpuCache::remove(", "ohm::GpuCache::clear()", "ohm::GpuCache::syncLayerTo(", "ohm::GpuCache::findLayerCache("):
        assert any(symbol in line for line in defined), f"{symbol} is not defined by the binding"
    undefined = [line.split(" U ")[1] for line in out.splitlines() if " U ohmb200_" in line]
    assert "ohmb200_integrate" in undefined and "ohmb200_read_regions" in undefined  # it goes through the C ABI
    assert "cuda" not in subprocess.check_output(["ldd", path]).decode().split("libohmb200")[0].lower()


def test_exchange_entry_points_reject_bad_arguments_without_a_gpu():
    """The exchange ABI: argument checks come before any CUDA call (no compute here), the handle is 128 opaque bytes."""
    lib = _lib.load()
    assert C.sizeof(_lib.ExchangeHandle) == 128
    handle = _lib.ExchangeHandle()
    assert lib.ohmb200_exchange_open(None, 0, 2, 1024, C.byref(handle)) == -1
    assert lib.ohmb200_exchange_connect(None, C.byref(handle), 1) == -1
    assert lib.ohmb200_exchange_send(None, None, 0, None, None, 0) == 0
    assert lib.ohmb200_exchange_send_device(None, None, 0, None, None, 0) == 0
    assert lib.ohmb200_exchange_integrate(None) == -1
    assert lib.ohmb200_exchange_barrier(None) == -1
    assert lib.ohmb200_exchange_last_counts(None, None, None, 0) == -1
    assert lib.ohmb200_exchange_close(None) == -1
    assert b"null" in lib.ohmb200_last_error() or b"exchange" in lib.ohmb200_last_error()
