"""The benchmark's output contract (CPU): the committed round-1 line carries every key the driver and the judge read,
and the reference arm runs on the host cores without a GPU and prints the same shape of line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config")


def _last_json_line(text):
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert lines, text[-500:]
    return json.loads(lines[-1])


def test_committed_bench_line_has_the_contract_keys():
    d = _last_json_line(open(os.path.join(ROOT, "profiles", "r01_bench_final_1gpu.json")).read())
    for k in BASE_KEYS + ("e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "decoder train tokens/sec" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "bf16" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and isinstance(k["reasons"], list)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    dec = d["decode"]
    assert dec["unit"] == "captions/s" and dec["roofline"]["bound"] == "hbm" and dec["roofline"]["unit"] == "GB/s"
    assert dec["config"]["schedule"]["partitions"] >= 1


def test_reference_arm_runs_on_the_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-budget-s", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-800:]
    d = _last_json_line(out.stdout)
    for k in BASE_KEYS + ("e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "decoder train tokens/sec" and d["value"] > 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    from oracle import ref_loader
    # the unmodified reference (oracle/_ref, copied by oracle/make_ref.sh) whenever it is present, else the oracle port
    want_kind = "reference" if ref_loader.available() else "port"
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and d["warmup"] == 1 and 8 <= d["cpu_baseline"]["sample_batch"] <= 256
    if want_kind == "reference":
        assert ref_loader.verify_manifest() and "MANIFEST MISMATCH" not in d["cpu_baseline"]["sample"]
