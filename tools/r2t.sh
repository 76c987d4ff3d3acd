mkdir -p gpurun_out/r2t
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t/pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2t/pytest.log
