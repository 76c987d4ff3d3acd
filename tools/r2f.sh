mkdir -p gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_p2p.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2f/pytest.log 2>&1; tail -5 gpurun_out/r2f/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f/bench_n1.json 2> gpurun_out/r2f/bench_n1.err; tail -2 gpurun_out/r2f/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f/bench_n1.json"))
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, "launches", d["gpu_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"p2p_|scan_reduce|scan_tiles|scan_apply" -c 90 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f/ncu.log 2>&1
python profiles/summarize.py launches gpurun_out/r2f/launches.csv | head -24
