"""Device map <-> `.ohm` file (ohm::save / ohm::load, ohm/MapSerialise.cpp): save_map gathers every layer of every region,
load_map recreates the map kind the file describes and uploads the blocks.  A loaded map must carry on exactly where
the saved one stopped — including the per-voxel bits the NDT and TSDF kernels keep beside the layers."""
import numpy as np
import pytest

import ohm_b200
from ohm_b200 import gpumap as gm
from ohm_b200 import ohmfile
from ohm_b200.lidar import cube_rays
from oracle import pyref as pr
from parity import compare_maps, make_pair

pytestmark = pytest.mark.gpu


def rays_for(seed, n=3000):
    rng = np.random.RandomState(seed)
    r = np.empty((2 * n, 3))
    r[0::2] = [0.05, 0.05, 0.05]
    r[1::2] = rng.uniform(-9, 9, size=(n, 3))
    return np.concatenate([cube_rays(3000), r])


def same_dump(a, b):
    assert sorted(a) == sorted(b)
    for key in a:
        assert sorted(a[key]) == sorted(b[key])
        for layer in a[key]:
            x, y = np.ascontiguousarray(a[key][layer]), np.ascontiguousarray(b[key][layer])
            assert np.array_equal(x.view(np.uint8).ravel(), y.view(np.uint8).ravel()), (key, gm.LAYER_NAMES[layer])


@pytest.mark.parametrize("mode", ["occupancy", "ndt", "ndt_tm", "tsdf"])
def test_save_load_continue(gpu, tmp_path, mode):
    kw = dict(layers=[gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_INCIDENT]) if mode == "occupancy" else {}
    g, c = make_pair(0.25, mode=mode, **kw)
    first, second = rays_for(1), rays_for(2)
    inten = np.linspace(0, 200, first.shape[0] // 2).astype(np.float32)
    g.integrate_rays(first, intensities=inten)
    c.integrate_rays(first, intensities=inten)
    path = tmp_path / "map.ohm"
    assert ohmfile.save_map(path, g) == g.region_count()
    loaded = ohmfile.load_map(path, device_bytes=1 << 30)
    assert type(loaded) is type(g) and loaded.layers() == g.layers()
    same_dump(g.dump(), loaded.dump())
    # the reference reads the same file (when oracle/_ref is on the box)
    if pr.available(build=False):
        same_dump(pr.ReferenceMap.load(path, g.layers()).dump(), g.dump())
    # carry on in the loaded map: same result as the CPU mapper that never stopped
    loaded.integrate_rays(second, intensities=inten)
    c.integrate_rays(second, intensities=inten)
    tol = {gm.LAYER_OCCUPANCY: (1e-5, 1e-5), gm.LAYER_INTENSITY: (1e-5, 1e-5)} if mode.startswith("ndt") else None
    compare_maps(loaded, c, tol_layers=tol)
    g.close()
    loaded.close()
