cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out/r2_final
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csarray_and_octree or find_neighbors or halos or sort_by_key or sfc_keys" > ${O}_sanitizer_memcheck.log 2>&1; echo "memcheck parity rc=$?"; tail -n 4 ${O}_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "reapply or halos_of_client" > ${O}_sanitizer_memcheck_domain.log 2>&1; echo "memcheck domain rc=$?"; tail -n 4 ${O}_sanitizer_memcheck_domain.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csarray_and_octree and u64d and uniform" > ${O}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 ${O}_sanitizer_racecheck.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
