mkdir -p gpurun_out/r2v
timeout 900 python -m pytest tests/test_gpu_sketch.py -m gpu -x -q -k "concurrent or overflow or multi_assembly or config2" > gpurun_out/r2v/pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2v/pytest.log
for ov in 1 0; do
MXE_SKETCH_OVERLAP=$ov python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v/bench_ov$ov.json 2> gpurun_out/r2v/bench_ov$ov.err; echo "bench rc=$?"; tail -2 gpurun_out/r2v/bench_ov$ov.err
done
python - <<'PY'
import json
for ov in (1,0):
    d=json.load(open(f"gpurun_out/r2v/bench_ov{ov}.json"))
    r=d["roofline"]
    print(ov, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), {k:round(v,2) for k,v in r["phase_ms_per_step"].items()}, "sketch_frac", round(r["sketch_frac"],4), "pack_cand", round(r["pack_cand_frac"],4), r["kernel"], round(r["frac"],3))
PY
