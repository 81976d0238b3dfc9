"""`.ohm` map files (ohm/MapSerialise.cpp, format version 0.5) to and from the device-resident map.

The file is what ohm::save() writes and ohm::load() reads:

    raw      header      marker 0x44330011, version {u32 major, u16 minor, u16 patch}, origin 3 x f64, region spatial
                         dimensions 3 x f64, region voxel dimensions 3 x i32, resolution, occupancy threshold, hit value,
                         miss value (f64), chunk count u32, first ray time f64, stamp u64, MapFlag u32
                                                                           (saveHeader, MapSerialise.cpp:287-325)
    raw      u32         number of MapInfo items                           (saveMapInfo, :253-284)
    zlib     items       name (u16 length + bytes), type u8, value         (saveItem, :95-250)
    zlib     layout      layers: name, flags, subsampling, voxel bytes, members {name, type, offset, clear value}
                                                                           (saveLayout, :328-370)
    zlib     chunks      region key 3 x i32, region centre 3 x f64, touched time f64, then per layer the touched stamp
                         u64 and the raw voxel block                       (saveChunk, :373-416)

Everything after the item count goes through one zlib stream (ohm/Stream.h:14 defines OHM_ZIP unconditionally;
OutputStream::write deflates, writeUncompressed does not).  A voxel block in the file is byte for byte a slab slot of
the device map (x + y*dx + z*dx*dy order, ohm/MapChunk.h:47-50), so saving is `ohmb200_read_regions` + deflate and
loading is inflate + `ohmb200_write_region`.

`write_ohm` / `read_ohm` work on plain dictionaries (no GPU needed: the CPU tests exchange files with the reference's
own ohm::save / ohm::load through oracle/_ref); `save_map` / `load_map` move a GpuMap to and from a file.
"""
import struct
import zlib

import numpy as np

from . import gpumap as gm

MARKER = 0x44330011
VERSION = (0, 5, 0)
_HEADER = "<IIHH3d3d3i4dIdQI"

# ohm/DataType.h:17-34
DT_UINT16, DT_UINT32, DT_FLOAT = 4, 6, 9
# ohm/MapFlag.h:16-38
FLAG_VOXEL_MEAN, FLAG_TRAVERSAL, FLAG_TOUCH_TIME, FLAG_INCIDENT, FLAG_TSDF, FLAG_SECONDARY = 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5, 1 << 6
_POS_INF_BITS = 0x7F800000  # unobservedOccupancyValue() as the occupancy member's clear value (DefaultLayer.cpp:87-91)

# layer -> (name, [(member, type, offset, clear value)])   ohm/DefaultLayer.cpp:29-330
LAYER_LAYOUT = {
    gm.LAYER_OCCUPANCY: ("occupancy", [("occupancy", DT_FLOAT, 0, _POS_INF_BITS)]),
    gm.LAYER_MEAN: ("mean", [("coord", DT_UINT32, 0, 0), ("count", DT_UINT32, 4, 0)]),
    gm.LAYER_TRAVERSAL: ("traversal", [("traversal", DT_FLOAT, 0, 0)]),
    gm.LAYER_TOUCH_TIME: ("touch_time", [("touch", DT_UINT32, 0, 0)]),
    gm.LAYER_INCIDENT: ("incident_normal", [("packed_normal", DT_UINT32, 0, 0)]),
    gm.LAYER_COVARIANCE: ("covariance", [("P00", DT_FLOAT, 0, 0), ("P01", DT_FLOAT, 4, 0), ("P11", DT_FLOAT, 8, 0),
                                         ("P02", DT_FLOAT, 12, 0), ("P12", DT_FLOAT, 16, 0), ("P22", DT_FLOAT, 20, 0)]),
    gm.LAYER_INTENSITY: ("intensity", [("mean", DT_FLOAT, 0, 0), ("cov", DT_FLOAT, 4, 0)]),
    gm.LAYER_HIT_MISS: ("hit_miss_count", [("hit_count", DT_UINT32, 0, 0), ("miss_count", DT_UINT32, 4, 0)]),
    gm.LAYER_TSDF: ("tsdf", [("weight", DT_FLOAT, 0, 0), ("distance", DT_FLOAT, 4, 0)]),
    gm.LAYER_SECONDARY: ("secondary_samples", [("m2", DT_FLOAT, 0, 0), ("range_mean", DT_UINT16, 4, 0),
                                               ("count", DT_UINT16, 6, 0)]),
}
LAYER_BY_NAME = {v[0]: k for k, v in LAYER_LAYOUT.items()}
_VOXEL_BYTES = {gm.LAYER_OCCUPANCY: 4, gm.LAYER_MEAN: 8, gm.LAYER_TRAVERSAL: 4, gm.LAYER_TOUCH_TIME: 4, gm.LAYER_INCIDENT: 4,
                gm.LAYER_COVARIANCE: 24, gm.LAYER_INTENSITY: 8, gm.LAYER_HIT_MISS: 8, gm.LAYER_TSDF: 8, gm.LAYER_SECONDARY: 8}
_TYPE_BYTES = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 8, 8: 8, 9: 4, 10: 8}

# MapValue types, ohm/MapInfo.h:38-53
_MV_FMT = {1: "<b", 2: "<B", 3: "<h", 4: "<H", 5: "<i", 6: "<I", 7: "<q", 8: "<Q", 9: "<f", 10: "<d", 11: "<B"}
MV_INT32, MV_UINT32, MV_FLOAT32, MV_FLOAT64, MV_BOOL, MV_STRING = 5, 6, 9, 10, 11, 12


class OhmFileError(Exception):
    pass


def map_flags(layers):
    flags = 0
    for layer, flag in ((gm.LAYER_MEAN, FLAG_VOXEL_MEAN), (gm.LAYER_TRAVERSAL, FLAG_TRAVERSAL),
                        (gm.LAYER_TOUCH_TIME, FLAG_TOUCH_TIME), (gm.LAYER_INCIDENT, FLAG_INCIDENT),
                        (gm.LAYER_SECONDARY, FLAG_SECONDARY)):
        flags |= flag if layer in layers else 0
    return flags


def _pack_items(info):
    out = bytearray()
    for name, (kind, value) in info.items():
        raw = name.encode()
        out += struct.pack("<H", len(raw)) + raw + struct.pack("<B", kind)
        if kind == MV_STRING:
            text = str(value).encode()
            out += struct.pack("<H", len(text)) + text
        else:
            out += struct.pack(_MV_FMT[kind], int(bool(value)) if kind == MV_BOOL else value)
    return bytes(out)


def write_ohm(path, header, regions, info=None, compress_level=-1):
    """Write a version 0.5 `.ohm` file.

    header: dict(resolution, origin (3), region_dim (3), threshold_value, hit_value, miss_value, first_ray_time, layers
            (layer ids in file order); optional stamp, flags)
    regions: {(rx, ry, rz): {layer: ndarray}} — each array holds one voxel block (any shape, voxel order x + y*dx + z*dx*dy)
    info: optional {name: (MapValue type, value)} MapInfo items
    """
    layers = list(header["layers"])
    dims = [int(d) for d in header["region_dim"]]
    res = float(header["resolution"])
    origin = [float(v) for v in header["origin"]]
    spatial = [res * d for d in dims]
    info = info or {}
    flags = int(header.get("flags", map_flags(layers)))
    raw = struct.pack(_HEADER, MARKER, VERSION[0], VERSION[1], VERSION[2], *origin, *spatial, *dims, res,
                      float(header["threshold_value"]), float(header["hit_value"]), float(header["miss_value"]),
                      len(regions), float(header.get("first_ray_time", -1.0)), int(header.get("stamp", 1)), flags)
    raw += struct.pack("<I", len(info))
    body = bytearray(_pack_items(info))
    body += struct.pack("<i", len(layers))
    for layer in layers:
        name, members = LAYER_LAYOUT[layer]
        enc = name.encode()
        body += struct.pack("<I", len(enc)) + enc + struct.pack("<IHII", 0, 0, _VOXEL_BYTES[layer], len(members))
        for member, kind, offset, clear in members:
            enc = member.encode()
            body += struct.pack("<I", len(enc)) + enc + struct.pack("<HHQ", kind, offset, clear)
    deflate = zlib.compressobj(compress_level)
    voxels = dims[0] * dims[1] * dims[2]
    with open(path, "wb") as f:
        f.write(raw)
        f.write(deflate.compress(bytes(body)))
        for key in sorted(regions):
            centre = [origin[a] + key[a] * spatial[a] for a in range(3)]   # MapRegion centre (ohm/MapRegion.cpp:24-30)
            chunk = bytearray(struct.pack("<3i3dd", int(key[0]), int(key[1]), int(key[2]), *centre, 0.0))
            for layer in layers:
                block = np.ascontiguousarray(regions[key][layer])
                if block.nbytes != voxels * _VOXEL_BYTES[layer]:
                    raise OhmFileError(f"layer {LAYER_LAYOUT[layer][0]} of region {key}: {block.nbytes} bytes, expected "
                                       f"{voxels * _VOXEL_BYTES[layer]}")
                chunk += struct.pack("<Q", 1)  # touched stamp of the layer
                chunk += block.tobytes()
            f.write(deflate.compress(bytes(chunk)))
        f.write(deflate.flush())


class _Reader:
    def __init__(self, data):
        self.data, self.at = data, 0

    def take(self, fmt):
        size = struct.calcsize(fmt)
        if self.at + size > len(self.data):
            raise OhmFileError("unexpected end of file")
        out = struct.unpack_from(fmt, self.data, self.at)
        self.at += size
        return out

    def raw(self, n):
        if self.at + n > len(self.data):
            raise OhmFileError("unexpected end of file")
        out = self.data[self.at:self.at + n]
        self.at += n
        return out


def read_ohm(path):
    """Read a `.ohm` file (format 0.4 / 0.5, compressed or not).  Returns (header, regions, info) as write_ohm takes them;
    header additionally carries `version`, `stamp`, `flags`, `layer_names` and `unknown_layers` (names this build has no
    slab for — their blocks are skipped)."""
    with open(path, "rb") as f:
        data = f.read()
    head = _Reader(data)
    marker, major = head.take("<II")
    if marker != MARKER:
        raise OhmFileError("not a versioned .ohm file (no header marker)")
    minor, patch = head.take("<HH")
    if (major, minor) not in ((0, 4), (0, 5)):
        raise OhmFileError(f"unsupported .ohm version {major}.{minor}.{patch}")
    origin = head.take("<3d")
    head.take("<3d")  # region spatial dimensions = resolution x voxel dimensions
    dims = head.take("<3i")
    res, threshold, hit, miss = head.take("<4d")
    (region_count,) = head.take("<I")
    first_ray_time = head.take("<d")[0] if minor >= 5 else -1.0
    (stamp,) = head.take("<Q")
    (flags,) = head.take("<I")
    (item_count,) = head.take("<I")
    rest = data[head.at:]
    if rest[:1] == b"\x78":  # zlib stream (deflateInit default header)
        try:
            rest = zlib.decompress(rest)
        except zlib.error as e:
            raise OhmFileError(f"corrupt compressed stream: {e}")
    r = _Reader(rest)
    info = {}
    for _ in range(item_count):
        (n,) = r.take("<H")
        name = r.raw(n).decode()
        (kind,) = r.take("<B")
        if kind == MV_STRING:
            (n,) = r.take("<H")
            info[name] = (kind, r.raw(n).decode())
        elif kind in _MV_FMT:
            info[name] = (kind, r.take(_MV_FMT[kind])[0])
        else:
            raise OhmFileError(f"unknown MapValue type {kind}")
    (layer_count,) = r.take("<i")
    file_layers = []  # (layer id or None, name, voxel bytes)
    for _ in range(layer_count):
        (n,) = r.take("<I")
        name = r.raw(n).decode()
        layer_flags, _, voxel_bytes, member_count = r.take("<IHII")
        members = []
        for _ in range(member_count):
            (n,) = r.take("<I")
            member = r.raw(n).decode()
            kind, offset, clear = r.take("<HHQ")
            members.append((member, kind, offset, clear))
        layer = LAYER_BY_NAME.get(name)
        if layer is not None:
            expect = [(m, k, o) for m, k, o, _ in LAYER_LAYOUT[layer][1]]
            if voxel_bytes != _VOXEL_BYTES[layer] or [(m, k, o) for m, k, o, _ in members] != expect:
                raise OhmFileError(f"layer {name}: voxel layout differs from the built-in one")
        file_layers.append((layer, name, voxel_bytes, layer_flags))
    voxels = dims[0] * dims[1] * dims[2]
    regions = {}
    for _ in range(region_count):
        key = r.take("<3i")
        r.take("<3dd")
        blocks = {}
        for layer, name, voxel_bytes, layer_flags in file_layers:
            if layer_flags & 1:
                continue  # MapLayer::kSkipSerialise: no stamp, no block in the file (MapSerialiseV0.4.cpp:60-65)
            r.take("<Q")
            block = r.raw(voxels * voxel_bytes)
            if layer is not None:
                dtype, width = gm.LAYER_DTYPES[layer]
                arr = np.frombuffer(block, dtype=dtype)
                blocks[layer] = arr.reshape(voxels, width) if width > 1 else arr
        regions[tuple(int(k) for k in key)] = blocks
    header = dict(resolution=res, origin=tuple(origin), region_dim=tuple(dims), threshold_value=threshold, hit_value=hit,
                  miss_value=miss, first_ray_time=first_ray_time, stamp=stamp, flags=flags, version=(major, minor, patch),
                  layers=[f[0] for f in file_layers if f[0] is not None], layer_names=[f[1] for f in file_layers],
                  unknown_layers=[f[1] for f in file_layers if f[0] is None])
    return header, regions, info


def _ndt_info(gpu):
    p = gpu.params
    mode = 2 if p.ndt_tm else 1
    return {  # NdtMap::updateMapInfo (ohm/NdtMap.cpp:178-192)
        "Ndt mode": (MV_INT32, mode),
        "Ndt mode name": (MV_STRING, "traversability" if p.ndt_tm else "occupancy"),
        "Ndt adaptation rate": (MV_FLOAT32, p.adaptation_rate),
        "Ndt sensor noise": (MV_FLOAT32, p.sensor_noise),
        "Ndt sample threshold": (MV_UINT32, p.sample_threshold),
        "Ndt reinitialisation threshold": (MV_FLOAT32, p.reinit_threshold),
        "Ndt reinitialisation point count": (MV_UINT32, p.reinit_count),
    }


def save_map(path, gpu, compress_level=-1):
    """ohm::save(path, map) for a device-resident map: sync, gather every layer of every region, write.  Returns the
    number of regions written."""
    gpu.sync_voxels()
    p = gpu.params
    layers = gpu.layers()
    header = dict(resolution=p.resolution, origin=tuple(p.origin), region_dim=tuple(p.region_dim),
                  threshold_value=p.threshold_value, hit_value=p.hit_value, miss_value=p.miss_value,
                  first_ray_time=gpu.first_ray_time(), layers=layers,
                  flags=map_flags(layers) | (FLAG_TSDF if gm.LAYER_TSDF in layers else 0))
    write_ohm(path, header, gpu.dump(), info=_ndt_info(gpu) if gpu.mode in ("ndt", "ndt_tm") else None,
              compress_level=compress_level)
    return gpu.region_count()


def load_map(path, device_bytes=0, device=0):
    """ohm::load(path, map) into a new device-resident map of the matching kind (GpuMap / GpuNdtMap / GpuTsdfMap).  The
    mapping parameters the file carries (resolution, origin, region dimensions, hit / miss / threshold, first ray time,
    the NDT items of the MapInfo block) are restored; the others keep their defaults."""
    header, regions, info = read_ohm(path)
    layers = header["layers"]
    kw = dict(device_bytes=device_bytes, device=device, origin=header["origin"], region_dim=header["region_dim"])
    if gm.LAYER_TSDF in layers:
        gpu = gm.GpuTsdfMap(header["resolution"], **kw)
    elif gm.LAYER_COVARIANCE in layers:
        gpu = gm.GpuNdtMap(header["resolution"], traversability=gm.LAYER_HIT_MISS in layers,
                           layers=[l for l in layers if l not in (gm.LAYER_COVARIANCE, gm.LAYER_INTENSITY, gm.LAYER_HIT_MISS)],
                           **kw)
    else:
        gpu = gm.GpuMap(header["resolution"], layers=layers, **kw)
    missing = [gm.LAYER_NAMES[l] for l in layers if l not in gpu.layers()]
    if missing:
        gpu.close()
        raise OhmFileError(f"the file holds layers this map kind does not: {missing}")
    params = dict(hit_value=header["hit_value"], miss_value=header["miss_value"], threshold_value=header["threshold_value"])
    for name, field in (("Ndt adaptation rate", "adaptation_rate"), ("Ndt sensor noise", "sensor_noise"),
                        ("Ndt sample threshold", "sample_threshold"), ("Ndt reinitialisation threshold", "reinit_threshold"),
                        ("Ndt reinitialisation point count", "reinit_count")):
        if name in info and gm.LAYER_COVARIANCE in layers:
            params[field] = info[name][1]
    gpu.set_params(**params)
    if header["first_ray_time"] >= 0:
        gpu.set_first_ray_time(header["first_ray_time"])
    for key, blocks in regions.items():
        for layer, block in blocks.items():
            gpu.write_region(key, layer, block)
    return gpu
