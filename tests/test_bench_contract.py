"""CPU: the reference arm of bench.py (`--impl reference`: the UNMODIFIED reference, imported through oracle/ref_shim.py from
/root/reference or from the vendored copy under baseline/_ref/, timed on the host cores; no GPU) runs and prints ONE JSON line
carrying the keys the driver's contract names, with the SAME `config` object our arm prints for that command line."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def _have_reference():
    from oracle import ref_shim
    return ref_shim.available()


def _check_line(d, metric, unit):
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["unit"] == d["unit"] and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(not _have_reference(), reason="no reference tree (/root/reference or baseline/_ref)")
def test_reference_arm_sim_line():
    d = _run("--workload", "sim", "--steps", "1", "--warmup", "0", "--ref-slice", "0.5")
    _check_line(d, "env_steps_per_sec", "env-steps/s")
    assert d["c_port"]["kind"] == "port" and d["c_port"]["value"] > d["value"]      # the C restatement beside the CPython reference
    assert 2600 < d["mean_ticks_per_env_step"] < 2900                               # 2669-2835 for this action


@pytest.mark.skipif(not _have_reference(), reason="no reference tree (/root/reference or baseline/_ref)")
def test_reference_arm_update_line_and_config_identity():
    import bench
    d = _run("--workload", "update", "--steps", "2", "--warmup", "1", "--batch", "256", "--replay", "20000")
    _check_line(d, "er_ddpg_updates_per_sec", "updates/s")
    args = types.SimpleNamespace(n_env=256, batch=256, replay=20000, actions="pid")
    assert d["config"] == bench.workload_config("update", args, 1)                   # byte-identical to what our arm prints


def test_both_arms_share_one_config_builder():
    """Every workload's `config` comes from bench.workload_config; arm-specific details live elsewhere (`timing`, `cpu_baseline`)."""
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(') >= 8 and '"config": {' not in src
    args = types.SimpleNamespace(n_env=65536, batch=4096, replay=1000000, actions="pid")
    for name in ("sim", "rollout", "update", "erddpg", "td3", "sac", "esac", "replay", "n1"):
        c = bench.workload_config(name, args, 8)
        assert isinstance(c["workload"], str) and "parallelism" in c


def test_vendored_reference_is_unmodified():
    from oracle import vendor_ref
    if not os.path.isdir(vendor_ref.DST):
        pytest.skip("baseline/_ref not populated (python oracle/vendor_ref.py)")
    assert vendor_ref.verify()
    if os.path.isdir(vendor_ref.SRC):       # build container: the copy equals the tree it was taken from
        for rel in vendor_ref._files(vendor_ref.SRC):
            assert vendor_ref._sha(os.path.join(vendor_ref.SRC, rel)) == vendor_ref._sha(os.path.join(vendor_ref.DST, rel)), rel
