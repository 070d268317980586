cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_exp_nb_bucket.*
CSB_TUNING=2=2 timeout 300 python tools/exp_neighbors.py --bucket 64 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
timeout 300 python tools/exp_neighbors.py --bucket 32 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
CSB_TUNING=2=2 timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "neighb or group" 2>&1 | tail -n 3
cut -c1-130 gpurun_out/r2_exp_nb_bucket.jsonl; tail -n 3 gpurun_out/r2_exp_nb_bucket.err
