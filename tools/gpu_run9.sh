cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"findNeighbors|nodeRange|leafContainment|groupBuild" --csv --log-file gpurun_out/r2_nbg_morton_launches.csv python tools/exp_neighbors.py 16777216 --config morton --only 0,0 --reps 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"findNeighbors|nodeRange|leafContainment|groupBuild" --csv --log-file gpurun_out/r2_nbg_uniform_launches.csv python tools/exp_neighbors.py --only 0,0 --reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighbors -c 1 -o gpurun_out/r2_nbg python tools/exp_neighbors.py --only 0,0 --reps 1 > gpurun_out/r2_ncu_nbg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighbors -c 1 -o gpurun_out/r2_nbg_morton python tools/exp_neighbors.py 16777216 --config morton --only 0,0 --reps 1 > gpurun_out/r2_ncu_nbg_m.log 2>&1
grep -v "^==" gpurun_out/r2_nbg_morton_launches.csv | cut -d, -f5,12- | tail -n 12
grep -v "^==" gpurun_out/r2_nbg_uniform_launches.csv | cut -d, -f5,12- | tail -n 12
