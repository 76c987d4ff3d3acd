set -x
mkdir -p gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_sketch.py -m gpu -x -q -k "messy or geometry or invalid or kw_sweep" > gpurun_out/r2c/pytest_sketch.log 2>&1
tail -3 gpurun_out/r2c/pytest_sketch.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c/bench_v4.json 2> gpurun_out/r2c/bench_v4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c/bench_v4.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["phase_ms_per_step"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pack2_kernel|scan_bs2" -c 2 -o gpurun_out/r2c/prof_front python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c/ncu.log 2>&1
tail -3 gpurun_out/r2c/ncu.log
