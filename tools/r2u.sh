mkdir -p gpurun_out/r2u
timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r2u/pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2u/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u/bench_n1.json 2> gpurun_out/r2u/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2u/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2u/bench_n1.json"))
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, d["gpu_launches"])
PY
