O=gpurun_out/final2; mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/final2/bench_n1.json")); r=d["roofline"]
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
print(r["kernel"], round(r["frac"],4), round(r["pack_cand_frac"],4), round(r["sketch_frac"],4))
PY
