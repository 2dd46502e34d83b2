"""CPU: the reference arm of bench.py (`--impl reference`) prints ONE JSON line with the keys the driver reads, runs the UNMODIFIED
reference when a copy is at hand (kind "reference") and the oracle port otherwise (kind "port"), and moves no bytes over PCIe."""
import json
import os
import subprocess
import sys
from helpers import ROOT as _R  # noqa: F401,E402  (sys.path: oracle/)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def _check(line, kinds):
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "images/s"
    assert line["metric"] == "images_per_sec_train_fwd_bwd_480x640_bf16" and line["value"] > 0 and line["steps"] == 1
    assert "BASELINE configs[2]" in line["config"]["workload"] and "sample" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in kinds and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_the_contract_line():
    import ref_shims  # noqa: F401  (tests/helpers puts oracle/ on sys.path)
    line = _run({})
    _check(line, ("reference",) if ref_shims.reference_available() else ("port",))


def test_reference_arm_falls_back_to_the_port_without_a_reference_copy():
    line = _run({"GWD_REFERENCE_ROOT": "/nonexistent"})
    _check(line, ("port",))
    assert "oracle/gwdepth_oracle.py" in line["cpu_baseline"]["sample"]
