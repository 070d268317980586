cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
CSB_TUNING=2=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighborsGroup -c 1 -o gpurun_out/r2_nbg_final_b64 python tools/exp_neighbors.py --bucket 64 --only 0,0 --reps 1 > gpurun_out/r2_ncu_nbg_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighborsGroup -c 1 -o gpurun_out/r2_nbg_final_b32 python tools/exp_neighbors.py --bucket 32 --only 0,0 --reps 1 >> gpurun_out/r2_ncu_nbg_final.log 2>&1
tail -n 2 gpurun_out/r2_ncu_nbg_final.log
