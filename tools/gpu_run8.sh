cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full.log 2>&1; tail -n 4 gpurun_out/r2_pytest_full.log
CSB_TUNING=2=2 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full_g.log 2>&1; tail -n 4 gpurun_out/r2_pytest_full_g.log
CSB_TUNING=2=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_full_l.log 2>&1; tail -n 4 gpurun_out/r2_pytest_full_l.log
