"""bench.py's JSON contract, checked where it can run without a GPU: the CPU reference arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_has_the_contract_keys():
    from oracle import refload
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "MLUPS" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == ("reference" if refload.available() else "port")
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    from lb_b200 import native
    if native.lib().lb_device_count() > 0:
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr


def test_every_evidence_file_named_in_the_profiles_readme_exists():
    """profiles/README.md is what the numbers in DESIGN.md / README.md point to: a file it names must be there."""
    import re
    prof = os.path.join(ROOT, "profiles")
    text = open(os.path.join(prof, "README.md")).read()
    names = set(re.findall(r"`((?:history/)?r[12]_[A-Za-z0-9_.\-]+\.(?:txt|json|csv))`", text))
    assert len(names) > 40
    missing = sorted(n for n in names if not os.path.exists(os.path.join(prof, n)))
    assert not missing, missing
    traffic = json.load(open(os.path.join(prof, "traffic.json")))
    for key in ("f32", "f64", "f32_march", "f64_march", "f32_march3", "f64_march3"):
        assert traffic[key]["dram_bytes_per_node_per_launch"] > 0 and len(traffic[key]["grid"]) == 2
        src = traffic[key]["source"].split(" ")[0]
        assert os.path.exists(os.path.join(ROOT, src)), src
