mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dropin.py tests/test_real_indexlr.py -m gpu -x -q > gpurun_out/r2e/pytest.log 2>&1; tail -5 gpurun_out/r2e/pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e/bench_n1.json 2> gpurun_out/r2e/bench_n1.err; tail -2 gpurun_out/r2e/bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e/bench_n1.json"))
r=d["roofline"]
print(d["value"], d["ms_per_step"], "e2e", d["e2e"])
print({k:r[k] for k in ("kernel","frac","pack_cand_frac","sketch_frac","launch_ms")}, r["kernels"])
print(d.get("cpu_baseline"))
PY
