"""`.ohm` files (ohm/MapSerialise.cpp, version 0.5) exchanged with the reference's own ohm::save / ohm::load.

CPU only: ohm_b200.ohmfile.write_ohm / read_ohm work on plain dictionaries, the maps come from the oracle and from
oracle/_ref.  (The GPU round trip — device map -> file -> device map — is tests/test_gpu_ohmfile.py.)
"""
import numpy as np
import pytest

from ohm_b200 import gpumap as gm
from ohm_b200 import ohmfile
from ohm_b200.lidar import cube_rays
from oracle import pyoracle as po
from oracle import pyref as pr

pytestmark = pytest.mark.skipif(not pr.available(), reason="oracle/_ref not built and /root/reference absent")


def same_regions(a, b):
    assert sorted(a) == sorted(b)
    for key in a:
        for layer in a[key]:
            x, y = np.ascontiguousarray(a[key][layer]), np.ascontiguousarray(b[key][layer])
            assert np.array_equal(x.view(np.uint8).ravel(), y.view(np.uint8).ravel()), (key, gm.LAYER_NAMES[layer])


def header_of(m, **extra):
    p = m.params
    h = dict(resolution=p.resolution, origin=tuple(p.origin), region_dim=tuple(p.region_dim), threshold_value=p.threshold_value,
             hit_value=p.hit_value, miss_value=p.miss_value, first_ray_time=-1.0, layers=m.layers())
    h.update(extra)
    return h


def rays_for(seed):
    rng = np.random.RandomState(seed)
    r = np.empty((2 * 1500, 3))
    r[0::2] = [0.05, 0.05, 0.05]
    r[1::2] = rng.uniform(-9, 9, size=(1500, 3))
    return np.concatenate([cube_rays(3000), r])


@pytest.mark.parametrize("mode,layers,kw", [
    ("occupancy", [gm.LAYER_OCCUPANCY], {}),
    ("occupancy", [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TOUCH_TIME, gm.LAYER_INCIDENT], dict(origin=(0.1, -0.2, 0.3))),
    ("occupancy", [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_TRAVERSAL], dict(region_dim=(16, 24, 8))),
    ("occupancy", [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_SECONDARY], {}),
    ("ndt", None, {}),
    ("ndt_tm", None, {}),
    ("tsdf", None, {}),
])
def test_files_we_write_load_in_the_reference(tmp_path, mode, layers, kw):
    """oracle map -> write_ohm -> ohm::load: header fields and every voxel block arrive unchanged."""
    okw = dict(kw)
    mode_layers = {"ndt": [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_COVARIANCE],
                   "ndt_tm": [gm.LAYER_OCCUPANCY, gm.LAYER_MEAN, gm.LAYER_COVARIANCE, gm.LAYER_INTENSITY, gm.LAYER_HIT_MISS],
                   "tsdf": [gm.LAYER_TSDF]}
    okw["layers"] = layers if layers is not None else mode_layers[mode]
    if mode == "ndt_tm":
        okw["ndt_tm"] = 1
    o = po.OracleMap(0.25, mode=mode, **okw)
    rays = rays_for(3)
    ts = np.linspace(10.0, 11.0, rays.shape[0] // 2) if gm.LAYER_TOUCH_TIME in o.layers() else None
    o.integrate_rays(rays, intensities=np.linspace(0, 200, rays.shape[0] // 2).astype(np.float32), timestamps=ts)
    if gm.LAYER_SECONDARY in o.layers():
        o.integrate_secondary(rays[:3000])
    path = tmp_path / "ours.ohm"
    first = o.first_ray_time() if ts is not None else -1.0
    ohmfile.write_ohm(path, header_of(o, first_ray_time=first), o.dump())
    r = pr.ReferenceMap.load(path, o.layers())
    same_regions(o.dump(), r.dump())
    assert r.header["resolution"] == o.params.resolution and r.header["region_dim"] == tuple(o.params.region_dim)
    assert r.header["origin"] == tuple(o.params.origin) and r.header["first_ray_time"] == first
    assert np.float32(r.header["hit_value"]) == o.params.hit_value and np.float32(r.header["miss_value"]) == o.params.miss_value


@pytest.mark.parametrize("mode", ["occupancy", "ndt", "tsdf"])
def test_files_the_reference_writes_load_here(tmp_path, mode):
    """reference map -> ohm::save (zlib-compressed) -> read_ohm: identical blocks, header and NDT map info."""
    kw = dict(layers=[gm.LAYER_OCCUPANCY, gm.LAYER_MEAN]) if mode == "occupancy" else {}
    r = pr.ReferenceMap(0.25, mode=mode, **kw)
    r.integrate_rays(rays_for(4))
    path = tmp_path / "theirs.ohm"
    r.save(path)
    header, regions, info = ohmfile.read_ohm(path)
    assert header["version"] == (0, 5, 0) and header["resolution"] == 0.25 and header["region_dim"] == (32, 32, 32)
    assert sorted(header["layers"]) == sorted(r.layers()) and not header["unknown_layers"]
    same_regions(r.dump(), regions)
    if mode == "ndt":
        assert info["Ndt sample threshold"][1] == 3 and abs(info["Ndt sensor noise"][1] - 0.05) < 1e-7
    # and back: what we read, written again, loads in the reference
    again = tmp_path / "again.ohm"
    ohmfile.write_ohm(again, header, regions, info=info)
    same_regions(pr.ReferenceMap.load(again, header["layers"]).dump(), r.dump())


def test_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.ohm"
    bad.write_bytes(b"not a map at all")
    with pytest.raises(ohmfile.OhmFileError):
        ohmfile.read_ohm(bad)
