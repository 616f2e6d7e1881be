cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_gpu_full.log
