cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k group_steered 2>&1 | tail -n 5
rm -f gpurun_out/r2_exp_nb_n4.jsonl
for t in "2=1" "2=2"; do
CSB_TUNING=$t timeout 300 python tools/exp_neighbors.py 33554432 --only 0,0 >> gpurun_out/r2_exp_nb_n4.jsonl 2>> gpurun_out/r2_exp_nb_n4.err
done
cat gpurun_out/r2_exp_nb_n4.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighborsGroup -c 1 -o gpurun_out/r2_nbg_b32 python tools/exp_neighbors.py --bucket 32 --only 0,0 --reps 1 > gpurun_out/r2_ncu_nbg_b32.log 2>&1
tail -n 2 gpurun_out/r2_ncu_nbg_b32.log
