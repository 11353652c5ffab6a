"""CPU checks of bench.py's contract pieces that do not need a GPU: the algorithmic-bytes formula of
SURVEY.md §8(d) / BASELINE.md §3, and the `--impl reference` arm (runs the CPU oracle on a bounded sample
and prints one JSON line with the required keys)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_baseline_table():
    # fwd+bwd bytes per image = 64*is^2 + F*(116 + 36*T2)
    for isz, F, T2, total in [(256, 1280, 36, 6001664), (512, 1280, 36, 18584576), (1024, 5120, 36, 74338304),
                              (64, 1280, 36, 2069504), (256, 1280, 1, 4388864)]:
        fwd, bwd = bench.alg_bytes_per_image(isz, F, T2)
        assert fwd + bwd == total
        assert fwd == F * (44 + 12 * T2) + 48 * isz * isz and bwd == 16 * isz * isz + F * (72 + 24 * T2)


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)   # no silent CPU fallback
