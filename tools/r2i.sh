mkdir -p gpurun_out/r2i
for f in tests/test_gpu_filter.py tests/test_gpu_p2p.py tests/test_gpu_dropin.py; do
  timeout 900 python -m pytest $f -m gpu -x -q > gpurun_out/r2i/pytest_$(basename $f).log 2>&1; echo "$f rc=$?"; tail -2 gpurun_out/r2i/pytest_$(basename $f).log
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i/bench_n1.json 2> gpurun_out/r2i/bench_n1.err; tail -2 gpurun_out/r2i/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2i/bench_n1.json"))
print(d["value"], d["ms_per_step"], {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, "launches", d["gpu_launches"])
PY
