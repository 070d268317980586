cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighborsKernel -c 1 -o gpurun_out/r2_nblegacy python tools/exp_neighbors.py --only 0,0 --reps 1 > gpurun_out/r2_ncu_nbcert.log 2>&1
tail -n 3 gpurun_out/r2_ncu_nbcert.log
