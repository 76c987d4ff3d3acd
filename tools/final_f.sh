O=gpurun_out/final3; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_filter.py tests/test_gpu_dropin.py tests/test_gpu_sketch.py -m gpu -x -q -k "not ipc" > $O/pytest_gpu_fast.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_fast.log
timeout 200 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -2 $O/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/final3/bench_n1.json")); r=d["roofline"]
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
print(r["kernel"], round(r["frac"],4), round(r["pack_cand_frac"],4), round(r["sketch_frac"],4), "t3", d.get("t3",{}).get("speedup_vs_t4"))
PY
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pack2|boundary2|scan_|dirty_fix|bitmap_|cand_|hash_pos|select_|empty_contig|gap_kernel|final_eval|p2p_|contig_bounds|record_start|pack_kernel" -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1; echo "ncu rc=$?"
for tool in initcheck memcheck; do
timeout 60 compute-sanitizer --tool $tool --target-processes all --log-file $O/sanitizer_$tool.txt python -m pytest tests/test_gpu_p2p.py tests/test_gpu_filter.py tests/test_gpu_sketch.py -m gpu -x -q -k "bucket_overflow or golden_steps23 or edge_cases or 1-1-100 or 3-2-250 or 2-1-100-0 or repeated or host_copy" > $O/sanitizer_${tool}_pytest.log 2>&1
echo "$tool rc=$?"; tail -1 $O/sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY" $O/sanitizer_$tool.txt | tail -1
done
timeout 100 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "full_bit_exact" > $O/pytest_fullsize.log 2>&1; echo "fullsize rc=$?"; tail -2 $O/pytest_fullsize.log
