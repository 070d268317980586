cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_exp_nb_bucket.*
for t in "2=2" "2=0"; do
CSB_TUNING=$t timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_nb_$t.log 2>&1; tail -n 4 gpurun_out/r2_pytest_nb_$t.log
done
for b in 64 32 16; do
timeout 300 python tools/exp_neighbors.py --bucket $b --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
done
timeout 300 python tools/exp_neighbors.py 16777216 --config morton --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
timeout 300 python tools/exp_neighbors.py --pbc 1 --bucket 32 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
cat gpurun_out/r2_exp_nb_bucket.jsonl; tail -n 3 gpurun_out/r2_exp_nb_bucket.err
