"""The CUDA path (through the C ABI) against the golden vectors of tests/golden/ — outputs of the reference's own CPU
mappers (tools/make_golden.py).  Same bars as the oracle-based parity tests: bit for bit, NDT log-odds 1e-5."""
import pytest

import ohm_b200
from golden_util import CASES, Golden
from ohm_b200 import gpumap as gm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_cuda_path_reproduces_the_reference_output(gpu, name):
    g = Golden(name)
    cls = {"occupancy": ohm_b200.GpuMap, "ndt": ohm_b200.GpuNdtMap, "ndt_tm": ohm_b200.GpuNdtMap,
           "tsdf": ohm_b200.GpuTsdfMap}[g.mode]
    kw = dict(g.params)
    if g.mode == "ndt_tm":
        kw["traversability"] = True
    m = cls(g.resolution, device_bytes=1 << 30, **kw)
    g.run(m)
    m.sync_voxels()
    tolerance = {}
    if g.mode in ("ndt", "ndt_tm"):
        tolerance[gm.LAYER_OCCUPANCY] = (1e-5, 1e-5)     # exp / log: glibc vs CUDA, and the order of the miss sum
        tolerance[gm.LAYER_INTENSITY] = (1e-5, 1e-5)
    if gm.LAYER_TRAVERSAL in m.layers():
        tolerance[gm.LAYER_TRAVERSAL] = (2e-5, 1e-6)     # fp32 running sum added in a different order
    g.compare(m.dump(), tolerance)
    m.close()
