O=gpurun_out/final3; mkdir -p $O
timeout 40 compute-sanitizer --tool initcheck --target-processes all --log-file $O/sanitizer_initcheck2.txt python -m pytest tests/test_gpu_p2p.py tests/test_gpu_filter.py tests/test_gpu_sketch.py -m gpu -q -k "bucket_overflow or golden_steps23 or edge_cases or 1-1-100 or 1-3-100 or host_copy" > $O/sanitizer_initcheck2_pytest.log 2>&1
echo "initcheck rc=$?"; tail -1 $O/sanitizer_initcheck2_pytest.log; grep -E "ERROR SUMMARY" $O/sanitizer_initcheck2.txt | tail -1
