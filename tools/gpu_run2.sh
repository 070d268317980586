set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q > gpurun_out/r2_pytest_configs.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_configs.log
CSB_TUNING=1=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_g1.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_g1.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_uniform.json 2> gpurun_out/r2_bench_uniform.err
timeout 900 python bench.py --config plummer --steps 3 --warmup 3 > gpurun_out/r2_bench_plummer.json 2> gpurun_out/r2_bench_plummer.err
timeout 900 python bench.py --config morton --steps 3 --warmup 3 > gpurun_out/r2_bench_morton.json 2> gpurun_out/r2_bench_morton.err
tail -n 5 gpurun_out/r2_pytest_configs.log gpurun_out/r2_pytest_g1.log
tail -n 5 gpurun_out/r2_bench_*.err
