mkdir -p gpurun_out/r2j
PYTHONFAULTHANDLER=1 timeout 900 python -X faulthandler -m pytest tests/test_gpu_dropin.py -m gpu -x -q --durations=8 > gpurun_out/r2j/pytest_dropin.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2j/pytest_dropin.log
