set -x
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_sketch.py -m gpu -x -q > gpurun_out/r2b/pytest_sketch.log 2>&1
tail -15 gpurun_out/r2b/pytest_sketch.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench_v4.json 2> gpurun_out/r2b/bench_v4.err
MXE_CAND_VARIANT=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench_v1.json 2> gpurun_out/r2b/bench_v1.err
python - <<'PY'
import json
for v in ("v4","v1"):
    try:
        d=json.load(open(f"gpurun_out/r2b/bench_{v}.json"))
        print(v, d["value"], d["ms_per_step"], d["roofline"]["phase_ms_per_step"])
    except Exception as e:
        print(v, "failed", e)
PY
