"""The oracle (oracle/ohm_oracle.c) against golden vectors: outputs of the REFERENCE's own CPU mappers and line walk,
generated in the build container by tools/make_golden.py from /root/reference and committed under tests/golden/.
These pin the oracle wherever the reference is not available (tests/test_oracle_vs_ref.py needs oracle/_ref)."""
import numpy as np
import pytest

from golden_util import CASES, GOLDEN, Golden
from oracle import pyoracle as po


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_output(name):
    g = Golden(name)
    kw = dict(g.params)
    if g.mode in ("ndt", "ndt_tm"):     # the oracle takes the resolved layer set (ohm::NdtMap adds these itself)
        kw["layers"] = [po.LAYER_OCCUPANCY, po.LAYER_MEAN, po.LAYER_COVARIANCE] + (
            [po.LAYER_INTENSITY, po.LAYER_HIT_MISS] if g.mode == "ndt_tm" else [])
        kw["ndt_tm"] = int(g.mode == "ndt_tm")
    if g.mode == "tsdf":
        kw["layers"] = [po.LAYER_TSDF]
    m = po.OracleMap(g.resolution, mode=g.mode, **kw)
    g.run(m)
    assert m.first_ray_time() == g.meta["first_ray_time"]
    g.compare(m.dump())          # every layer of every region, bit for bit (NDT log-odds included: same libm)
    m.close()


def test_oracle_line_walk_matches_the_reference():
    import os
    z = np.load(os.path.join(GOLDEN, "linewalk.npz"))
    m = po.OracleMap(0.25)
    starts, ends = z["starts"], z["ends"]
    visits = 0
    for i in range(len(starts)):
        for flags in (0, 1, 2, 3):
            keys, enter, exit_ = m.walk_segment(starts[i], ends[i], flags)
            assert np.array_equal(keys, z[f"keys_{i}_{flags}"])
            assert np.array_equal(enter.view(np.uint64), z[f"enter_{i}_{flags}"].view(np.uint64))
            assert np.array_equal(exit_.view(np.uint64), z[f"exit_{i}_{flags}"].view(np.uint64))
            visits += len(keys)
    assert visits > 10000
    m.close()
