set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt
free -g >> gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v0.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_v0.log
CSB_TUNING=0=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_v1.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_v1.log
CSB_TUNING=0=1,1=1 timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_v11.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_v11.log
timeout 300 python tools/exp_neighbors.py > gpurun_out/r2_exp_nb_uniform.jsonl 2> gpurun_out/r2_exp_nb_uniform.err
timeout 300 python tools/exp_neighbors.py 16777216 --config morton > gpurun_out/r2_exp_nb_morton.jsonl 2> gpurun_out/r2_exp_nb_morton.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighbors -c 1 -o gpurun_out/r2_nb_v11 python tools/exp_neighbors.py --only 1,1 --reps 1 > gpurun_out/r2_ncu_nb.log 2>&1
tail -3 gpurun_out/r2_pytest_v0.log gpurun_out/r2_pytest_v1.log gpurun_out/r2_pytest_v11.log
cat gpurun_out/r2_exp_nb_uniform.jsonl gpurun_out/r2_exp_nb_morton.jsonl
