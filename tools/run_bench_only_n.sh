cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"]["neighbors"]["ms"], d["stages"]["domain_sync_cold"]["ms"], d["stages"]["domain_sync_steady"]["ms"], d["checks"]["transport_parity"]["bit_identical_vs_reference"])
PY
