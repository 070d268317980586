cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_exp_nb_bucket.*
timeout 300 python tools/exp_neighbors.py --bucket 64 --pbc 1 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
cut -c1-130 gpurun_out/r2_exp_nb_bucket.jsonl; tail -n 3 gpurun_out/r2_exp_nb_bucket.err
