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
    out = run_bench({"OMP_NUM_THREADS": "1"}, "--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "65536")
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
    # the reference arm runs the workload our arm prints as `config` (N=1: at full size) and publishes the digests
    # of its results so that the two lines can be compared
    assert d["config"]["particles_per_gpu"] == 65536 and cb["particles_per_rank"] == 65536
    assert d["results"]["digest"]["num_leaves"] > 0 and d["results"]["digest"]["nc_sum"] > 0


def test_reference_arm_other_workloads_and_ranks():
    for extra in (["--config", "plummer"], ["--config", "morton"]):
        out = run_bench({}, "--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "32768", *extra)
        d = json.loads(out.strip())
        assert d["config"]["name"] == extra[1] and d["value"] > 0
    # N>1: a reduced-size run of ONE reference Domain over N ranks (threads), not N single-rank runs
    out = run_bench({}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-n", "262144")
    d = json.loads(out.strip())
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["ranks"] == 2 and d["cpu_baseline"]["particles_per_rank"] == 32768


def test_reference_arm_is_silent_on_other_ranks():
    out = run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2",
                    "--steps", "1", "--warmup", "0", "--ref-n", "65536")
    assert out.strip() == ""
