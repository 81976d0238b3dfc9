"""Loader of tests/golden/*.npz — outputs of the reference itself (tools/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["config1_occupancy", "occupancy_layers_flags", "occupancy_clip_filter", "ndt_tm", "tsdf"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.mode = self.meta["mode"]
        self.resolution = self.meta["resolution"]
        self.params = dict(self.meta["params"])
        for k in ("region_dim", "origin"):
            if k in self.params:
                self.params[k] = tuple(self.params[k])
        self.passes = []
        for i in range(self.meta["passes"]):
            self.passes.append((z[f"pass{i}_rays"],
                                z[f"pass{i}_intensities"] if f"pass{i}_intensities" in z.files else None,
                                z[f"pass{i}_timestamps"] if f"pass{i}_timestamps" in z.files else None,
                                int(z[f"pass{i}_flags"][0])))
        self.regions = {}
        for f in z.files:
            if f.startswith("region_"):
                head, layer = f.rsplit("_layer", 1)
                key = tuple(int(v) for v in head[len("region_"):].split("_"))
                self.regions.setdefault(key, {})[int(layer)] = z[f]

    def run(self, m):
        for rays, intensities, timestamps, flags in self.passes:
            m.integrate_rays(rays, intensities, timestamps, flags)

    def compare(self, dump, tolerance=None):
        """dump = {(rx,ry,rz): {layer: array}}; tolerance = {layer: (rtol, atol)} for the layers not held bit for bit."""
        assert sorted(dump.keys()) == sorted(self.regions.keys()), "region sets differ"
        for key, layers in self.regions.items():
            for layer, want in layers.items():
                got = np.ascontiguousarray(dump[key][layer]).reshape(want.shape)
                if tolerance and layer in tolerance:
                    rtol, atol = tolerance[layer]
                    g = np.nan_to_num(got.astype(np.float64), posinf=1e30, neginf=-1e30)
                    w = np.nan_to_num(want.astype(np.float64), posinf=1e30, neginf=-1e30)
                    assert np.all(np.abs(g - w) <= atol + rtol * np.abs(w)), f"region {key} layer {layer} beyond tolerance"
                else:
                    assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), (
                        f"region {key} layer {layer} differs from the reference's output")
