"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU implementation of the path,
compiled under oracle/_ref, or the oracle port when that is absent) prints the agreed JSON line, and the B200 arm
refuses to run without a device instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["metric"].startswith("megapixel frame-pairs/sec") and line["unit"] == "Mpx-pairs/s"
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in json.loads(r.stdout.strip().splitlines()[-1])["error"]


def test_algorithmic_bytes_match_the_survey_figures():
    """the per-unit figures the roofline is quoted on.  configs[1]: 43,027,200 B per 1080p pair (level-0
    kernel alone 24,883,200 B); configs[0]: 3,072,000 B; configs[2]: 179,884,800 B; configs[4] with 4 levels: 719,539,200 B."""
    sys.path.insert(0, ROOT)
    import bench

    def pair_bytes(w, h, levels):
        lk = sum(bench.level_bytes(w, h, k, levels, 1, 0 < k < levels - 1) for k in range(levels))
        pyr = 2 * sum((w >> (k - 1)) * (h >> (k - 1)) + (w >> k) * (h >> k) for k in range(1, levels))
        return lk + pyr

    assert bench.level_bytes(1920, 1080, 0, 3, 1, False) == 24_883_200
    assert pair_bytes(1920, 1080, 3) == 43_027_200
    assert pair_bytes(640, 480, 1) == 3_072_000
    assert pair_bytes(3840, 2160, 4) == 179_884_800
    assert pair_bytes(7680, 4320, 4) == 719_539_200


def test_both_arms_print_the_same_config():
    """The driver compares the arms' `config`: it is one constant, and everything run-specific lives under `run`."""
    sys.path.insert(0, ROOT)
    import bench

    assert set(bench.CONFIG) == {"workload"} and "1920x1080" in bench.CONFIG["workload"]
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["config"] == bench.CONFIG and "pairs_per_step" in line["run"]
