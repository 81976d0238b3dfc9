"""Host-side pieces of bench.py that need no GPU: the byte models of SURVEY.md §8d and the rule that ncu traffic is only
quoted for the build it was measured on."""
import json
import os

import bench


def test_algorithmic_byte_models_follow_survey_8d():
    v, s, n = 27_615_730, 80_000, 131_072
    assert bench.algorithmic_bytes(2, v, s, n) == 8 * v + 44 * n == 226_693_008
    assert bench.algorithmic_bytes(3, v, s, n) == (8 + 8 + 24) * v + (8 + 16 + 48) * s + 44 * n
    assert bench.algorithmic_bytes(4, v, s, n) == 16 * v + 68 * n


def test_traffic_is_quoted_only_for_the_sources_it_was_measured_on(tmp_path, monkeypatch):
    real = os.path.join(bench.ROOT, "profiles", "traffic.json")
    t = json.load(open(real))
    assert set(t["kernels"]) >= {"walkRegions", "walkRegionsNdt", "walkRegionsTsdf"}
    current = bench.source_hash()
    got, how = bench.ncu_traffic("walkRegions")
    if t["source_hash"] == current:
        assert got == t["kernels"]["walkRegions"]["dram_bytes_per_launch"] and got > 10_000_000
    else:
        assert got is None and "another build" in how
    # a capture of some other build is never quoted
    monkeypatch.setattr(bench, "source_hash", lambda: "0" * 16)
    got, how = bench.ncu_traffic("walkRegions")
    assert got is None and "another build" in how


def test_every_config_names_its_workload_and_model():
    for number, cfg in bench.CONFIGS.items():
        assert cfg["mode"] in ("occupancy", "ndt", "tsdf") and cfg["resolution"] in (0.1, 0.05)
        assert "SURVEY" in cfg["model"] and cfg["workload"]
    assert bench.own_sweep_index(3, 5, 8) == 29
