cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for t in "2=2" "2=1"; do
CSB_TUNING=$t timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"findNeighbors" --csv --log-file gpurun_out/r2_nb_pbc_$t.csv python tools/exp_neighbors.py --pbc 1 --bucket 32 --only 0,0 --reps 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_nb_pbc_$t.csv | awk -F'","' '{print substr($5,1,60), $(NF-2), $NF}' | tail -n 8
done
