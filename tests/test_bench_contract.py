"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the keys the driver
reads, uses the host cores even when the launcher exports OMP_NUM_THREADS=1 (torch.distributed.run does), and stays
silent on the ranks other than 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line():
    out = run_bench({"OMP_NUM_THREADS": "1"}, "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-n", "65536")
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mparticles/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Mparticles/s, Domain::sync + findNeighbors")
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["sample"]
    if cb["kind"] == "reference":
        assert cb["cores"] == os.cpu_count(), "the reference arm must not inherit OMP_NUM_THREADS=1 from the launcher"


def test_reference_arm_is_silent_on_other_ranks():
    out = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2",
                    "--steps", "1", "--warmup", "0", "--ref-n", "65536")
    assert out.strip() == ""
