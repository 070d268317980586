set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v6.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_v6.log
tail -n 30 gpurun_out/r2_pytest_v6.log
timeout 300 python tools/exp_neighbors.py > gpurun_out/r2_exp_nb_uniform.jsonl 2> gpurun_out/r2_exp_nb_uniform.err
timeout 300 python tools/exp_neighbors.py 16777216 --config morton > gpurun_out/r2_exp_nb_morton.jsonl 2> gpurun_out/r2_exp_nb_morton.err
cat gpurun_out/r2_exp_nb_uniform.jsonl gpurun_out/r2_exp_nb_morton.jsonl
tail -n 5 gpurun_out/r2_exp_nb_uniform.err gpurun_out/r2_exp_nb_morton.err
