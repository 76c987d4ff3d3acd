mkdir -p gpurun_out/r2q
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --log-file gpurun_out/r2q/sanitizer_$tool.txt python -m pytest tests/test_gpu_sketch.py tests/test_gpu_p2p.py tests/test_gpu_filter.py -m gpu -x -q -k "messy or edge_cases or geometry or invalid_bytes or lockstep_ranks or golden_steps23 or multi_assembly" > gpurun_out/r2q/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/r2q/sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2q/sanitizer_$tool.txt | tail -2
done
