cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo rc=$? >> gpurun_out/r2_final_pytest.log; tail -n 3 gpurun_out/r2_final_pytest.log
timeout 600 python bench.py > gpurun_out/r2_final_bench_uniform_c.json 2> gpurun_out/r2_final_bench_uniform_c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_final_bench_uniform_c.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k:(v.get("ms"),v.get("frac")) for k,v in d["stages"].items()}, d["checks"]["sample_bit_identical_vs_reference"]["identical"], d["gpu_launches"])
PY
