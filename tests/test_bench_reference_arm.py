"""`bench.py --impl reference` (the CPU arm of the bench contract) runs without a GPU and emits the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "prefixes/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "prefixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("raft5 depth-50")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_k1_profile_describes_the_current_kernel_sources():
    """bench.py reports profile-derived figures (DRAM traffic, issue-slot utilisation) only from a capture of the
    kernel sources it was built from; a stale committed profile fails here instead of going unnoticed."""
    from demi_b200 import build
    prof = json.load(open(os.path.join(ROOT, "profiles", "k1_profile.json")))
    assert prof["k1_id"] == build.k1_source_id(), "re-run tools/profile_k1.py on the GPU box and commit profiles/k1_profile.json"
    assert prof["dram_bytes_per_prefix"] > 0 and 0 < prof["issue_slot_utilisation"] <= 1
