cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -n 3; fi
CSB_TUNING=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n${N}_t.json 2> gpurun_out/r2_bench_n${N}_t.err
tail -n 2 gpurun_out/r2_bench_n${N}_t.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_t.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stages"]["neighbors"]["ms"], d["stages"]["domain_sync_cold"]["ms"], d["stages"]["domain_sync_steady"]["ms"], d["results"]["focus_leaves"], d["checks"]["transport_parity"]["bit_identical_vs_reference"])
PY
CSB_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 1 --warmup 3 --parity-n 65536 > gpurun_out/r2_trace_n$N.json 2> gpurun_out/r2_trace_n$N.err
grep "csb rank 0" gpurun_out/r2_trace_n$N.err | grep -E "exchangeParticles|gatherArrays|keys\+sort received" | tail -n 6
