mkdir -p gpurun_out/r2d
for tau in 5 6 7 8 9 11; do
  MXE_TAU=$tau python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2d/bench_tau$tau.json 2> gpurun_out/r2d/bench_tau$tau.err
done
python - <<'PY'
import json
for t in (5,6,7,8,9,11):
    try:
        d=json.load(open(f"gpurun_out/r2d/bench_tau{t}.json"))
        print(t, round(d["value"],1), round(d["ms_per_step"],3), {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()})
    except Exception as e:
        print(t, "failed", e)
PY
