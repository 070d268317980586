# round-2 evidence: tests, bench lines of the three workloads, reference arm, ncu launch list and --set full captures.
# The .ncu-rep files are reduced to text on the box (gpurun brings back at most 64 MiB).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2_final
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > ${O}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo rc=$? >> ${O}_pytest.log
timeout 300 python __graft_entry__.py --smoke > ${O}_smoke.log 2>&1
timeout 900 python bench.py > ${O}_bench_uniform.json 2> ${O}_bench_uniform.err
timeout 900 python bench.py --config plummer > ${O}_bench_plummer.json 2> ${O}_bench_plummer.err
timeout 900 python bench.py --config morton > ${O}_bench_morton.json 2> ${O}_bench_morton.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > ${O}_bench_reference.json 2> ${O}_bench_reference.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_one_step.csv python tools/profile_stages.py > ${O}_ncu1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${O}_launches_bench.csv python bench.py --steps 2 --warmup 1 --ref-n 1048576 > ${O}_ncu2.log 2>&1
K='regex:onesweep|radixHistogram|sfcKeys|gatherRec4|packRec4|nodeCountsPipelined|coarseBounds|findNeighborsKernel|findHalos|boundingBox|geoCenters|linkTree|nodeOps'
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$K" -o /tmp/full python tools/profile_stages.py > ${O}_ncu3.log 2>&1
ncu -i /tmp/full.ncu-rep --page raw --csv > /tmp/full_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/full_raw.csv > ${O}_full_summary.txt
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:findNeighborsKernel > /tmp/nb_src.csv 2>/dev/null
python tools/ncu_regions.py /tmp/nb_src.csv > ${O}_neighbors_regions_lanes.txt 2>&1
ncu -i /tmp/full.ncu-rep --page source --csv --kernel-name regex:onesweep --launch-skip 0 --launch-count 1 > /tmp/os_src.csv 2>/dev/null
python tools/ncu_regions.py /tmp/os_src.csv > ${O}_onesweep_regions.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findNeighborsGroup -c 1 -o /tmp/grp python tools/exp_neighbors.py --bucket 32 --only 0,0 --reps 1 > ${O}_ncu4.log 2>&1
ncu -i /tmp/grp.ncu-rep --page raw --csv > /tmp/grp_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/grp_raw.csv > ${O}_group_b32_summary.txt
ncu -i /tmp/grp.ncu-rep --page source --csv > /tmp/grp_src.csv 2>/dev/null
python tools/ncu_regions.py /tmp/grp_src.csv > ${O}_neighbors_regions_group_b32.txt 2>&1
tail -n 3 ${O}_pytest.log ${O}_smoke.log
du -sh gpurun_out
ls -la gpurun_out | grep r2_final
