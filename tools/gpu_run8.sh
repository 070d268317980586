cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_exp_nb_bucket.*
CSB_TUNING=2=2 timeout 900 python -m pytest tests -m gpu -x -q -k "neighb or bit_identical or domain or multirank" > gpurun_out/r2_pytest_nb_g.log 2>&1; tail -n 4 gpurun_out/r2_pytest_nb_g.log
CSB_TUNING=2=2 timeout 300 python tools/exp_neighbors.py --bucket 64 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
CSB_TUNING=2=2 timeout 300 python tools/exp_neighbors.py 33554432 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
for b in 32 16; do
timeout 300 python tools/exp_neighbors.py --bucket $b --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
done
timeout 300 python tools/exp_neighbors.py --pbc 1 --bucket 32 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
cat gpurun_out/r2_exp_nb_bucket.jsonl; tail -n 3 gpurun_out/r2_exp_nb_bucket.err
