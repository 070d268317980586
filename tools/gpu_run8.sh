cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/r2_exp_nb_bucket.*
for c in 64 128 256; do
echo "coarse $c" >> gpurun_out/r2_exp_nb_bucket.jsonl
CSB_TUNING=2=2,3=$c timeout 300 python tools/exp_neighbors.py 33554432 --pbc 1 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
CSB_TUNING=2=2,3=$c timeout 300 python tools/exp_neighbors.py --bucket 32 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
CSB_TUNING=2=2,3=$c timeout 300 python tools/exp_neighbors.py --bucket 64 --only 0,0 >> gpurun_out/r2_exp_nb_bucket.jsonl 2>> gpurun_out/r2_exp_nb_bucket.err
done
cut -c1-130 gpurun_out/r2_exp_nb_bucket.jsonl; tail -n 3 gpurun_out/r2_exp_nb_bucket.err
