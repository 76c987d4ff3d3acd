mkdir -p gpurun_out/r2g
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g/pytest.log 2>&1; tail -4 gpurun_out/r2g/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g/bench_n1.json 2> gpurun_out/r2g/bench_n1.err; tail -2 gpurun_out/r2g/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g/bench_n1.json"))
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, "launches", d["gpu_launches"])
PY
