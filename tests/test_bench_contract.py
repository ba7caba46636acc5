"""CPU: the bench line committed under profiles/ (printed by bench.py on a B200) carries every key the driver's
contract names, with consistent values -- guards bench.py's JSON against accidental edits."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(next(l for l in f if l.startswith("{")))


def test_bench_line_has_the_contract_keys():
    d = _load("r2n_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"].replace("x", "×") == base["metric"].replace("x", "×")
    assert d["config"]["workload"].startswith("bs=32 synthetic 384x384 greedy decode on 1xB200")
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    # value is whole-job throughput over the timed steps
    imgs = d["config"]["global_batch"] * d["steps"]
    assert abs(d["value"] - imgs / (d["ms_per_step"] * d["steps"] / 1000.0)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 32 * 3 * 384 * 384 * 4 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    ck = d["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_and_two_gpu_lines():
    ref = _load("r2k_bench_ref.json")
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["value"] == ref["value"] and ref["unit"] == "images/s" and ref["cpu_baseline"]["kind"] == "reference"
    one, two = _load("r2n_bench.json"), _load("r2i_bench_2gpu.json")
    assert two["n_gpus"] == 2 and two["config"]["global_batch"] == 64
    assert 1.8 < two["value"] / one["value"] < 2.1          # weak scaling: images are independent


def test_lines_of_the_other_baseline_configs():
    """BASELINE.json configs 1, 3, 4, 5 (`bench.py --config cK`): same contract keys, the workload names its config."""
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    want = {"c1": "single 384x384 image", "c3": "bs=256 384x384 beam_size=5", "c4": "bs=2048 384x384 greedy", "c5": "bs=8 1024x1024"}
    for name, prefix in want.items():
        d = _load(f"r2l_bench_{name}.json" if name == "c1" else f"r2k_bench_{name}.json")
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                  "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert k in d, (name, k)
        assert d["metric"].replace("x", "×") == base["metric"].replace("x", "×")
        assert d["config"]["workload"].startswith(prefix) and d["config"]["name"] == name
        assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0 and d["warmup"] >= 3
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
