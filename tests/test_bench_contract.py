"""CPU: the bench line committed under profiles/ (printed by `python bench.py` on a B200, never under a profiler) carries
every key of the driver's contract, with consistent numbers: value = bases / time, roofline.frac = achieved / peak,
achieved = algorithmic bytes / launch time, e2e counted from the copied tensors."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    for l in open(os.path.join(ROOT, "profiles", name)):
        if l.startswith("{"):
            return json.loads(l)
    raise AssertionError(name)


def test_single_gpu_line_has_the_contract_keys():
    d = _line("r02_d_bench_configs2_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["unit"] == "Gbases/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    bases = d["config"]["bases_per_step"]
    assert abs(d["value"] - bases / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == bases and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                                   # the end-to-end number includes the PCIe copies
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    n_mx = sum(d["config"]["minimizers"])
    assert abs(r["algorithmic_bytes_per_launch"] - (bases + 16 * n_mx) / 2) < 1.0     # 1 B per base + 16 B per minimizer, per assembly
    assert r["traffic"] is None or r["traffic"] > 0
    assert 0 < r["sketch_frac"] < r["pack_cand_frac"] < r["frac"] < 1
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_two_gpu_line_carries_parity():
    d = _line("r02_d_bench_configs2_2gpu_p2p.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "strong"
    p = d["parity"]
    assert p["vs_single_gpu"] is True and p["mismatch"] == [] and p["digest"] == p["single_gpu_digest"]
