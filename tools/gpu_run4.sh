set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nbSearch -c 1 -o gpurun_out/r2_nbsearch python tools/exp_neighbors.py --only 2,0 --reps 1 > gpurun_out/r2_ncu_nbsearch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nb|groupBuild|findNeighbors" --csv --log-file gpurun_out/r2_nb_launches.csv python tools/exp_neighbors.py --only 2,0 --reps 1 > /dev/null 2>&1
cat gpurun_out/r2_nb_launches.csv | tail -8
