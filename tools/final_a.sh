set -x
O=gpurun_out/final; mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 5 --warmup 3 --with-n --no-cpu-baseline > $O/bench_n1_withN.json 2> $O/bench_n1_withN.err
python bench.py --steps 3 --warmup 2 --workload c4 --no-cpu-baseline > $O/bench_n1_c4.json 2> $O/bench_n1_c4.err
python bench.py --steps 2 --warmup 1 --sweep --no-cpu-baseline > $O/sweep_n1.json 2> $O/sweep_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pack2_kernel|scan_bs2|p2p_bucket|select_kernel|cand_hash_pos|cand_extract|final_eval|p2p_edge_emit|p2p_scatter_kernel" -s 18 -c 9 -o $O/prof_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
for tool in memcheck racecheck synccheck initcheck; do
  sel="messy or edge_cases or geometry or invalid_bytes or golden_steps23 or multi_assembly or overflow or concurrent or 1-1-100"
  if [ $tool = racecheck ]; then sel="$sel or lockstep_ranks"; fi
  timeout 900 compute-sanitizer --tool $tool --target-processes all --log-file $O/sanitizer_$tool.txt python -m pytest tests/test_gpu_sketch.py tests/test_gpu_p2p.py tests/test_gpu_filter.py -m gpu -x -q -k "$sel" > $O/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -1 $O/sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/sanitizer_$tool.txt | tail -1
done
