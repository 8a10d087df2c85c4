"""bench.py's reference arm runs on the CPU (the reference's shaders compiled as C++, or the oracle port, with all host threads): its JSON line can be checked here.
The b200 arm needs a GPU; its line is recorded under profiles/ and checked for the same keys."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "cpu_baseline", "e2e"}


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-1500:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    # "reference" = the reference's own GLSL compiled as C++ (oracle/_ref/libglsl_ref.so, built where /root/reference exists and
    # shipped to the GPU box), "port" = the restatement, only when that library is absent
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libglsl_ref.so"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_recorded_b200_lines_carry_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_v1[01]_bench_n*.json")))
    assert files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        need = BASE_KEYS | {"roofline", "gpu_launches", "clocks"}
        if d["n_gpus"] > 1:
            need = need - {"cpu_baseline"}        # the CPU baseline is timed on rank 0 at N = 1 only
        assert need <= set(d), f
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
