"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line on stdout with the
keys the driver reads, on the same workload / metric / unit as the CUDA arm, and the CUDA arm refuses to run
without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hypotheses_per_sec" and d["unit"] == "hypotheses/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    if have_ref:        # the unmodified reference, hashed against its manifest
        assert d["cpu_baseline"]["reference"]["verified"] == "sha256 of every file equals MANIFEST.json"
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"]["workload"] == bench.WORKLOAD_NAME


def test_reference_arm_of_the_other_configs():
    """cfg1 (Stewenius loop body) and a training config through the reference's own entry points."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")):
        return
    for cfg in ("cfg1", "cfg3"):
        p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--config", cfg)
        assert p.returncode == 0, p.stderr[-2000:]
        d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
        assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["workload"].startswith(cfg)
        assert d["cpu_baseline"]["kind"] == "reference"


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env,
                       timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_needs_a_device():
    import torch

    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_tensor_side_arithmetic():
    """The tcgen05 figures bench.py adds for the tensor-core scorer: pure arithmetic, checked on the cfg2 numbers
    measured on the B200 (138 910 models x 2000 correspondences in 0.122 ms)."""
    import bench

    t = bench.tensor_side(138910, 2000, 0.122, "tc_tf32")
    tiles = 16 * 1086
    assert abs(t["tensor_tflops"] - tiles * 6 * 2 * 128 * 256 * 8 / 0.122e-3 / 1e12) < 1e-6
    assert 0.1 < t["tensor_frac"] < 1.0
    b = bench.tensor_side(138910, 2000, 0.122, "tc_bf16")
    assert abs(b["tensor_tflops"] / t["tensor_tflops"] - 2.0) < 1e-9 and b["tensor_peak_tflops"] == 2 * t["tensor_peak_tflops"]
    assert "unavailable" in bench.tensor_side(1, 1, 0.0, "tc_tf32")["tensor_note"]
