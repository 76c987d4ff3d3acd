O=gpurun_out/r2y; mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_p2p.py tests/test_gpu_filter.py -m gpu -x -q --durations=6 > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/bench_$name.err
}
run base MXE_NOP=1
run ov2 MXE_SKETCH_OVERLAP=2
python - <<'PY'
import json
for nm in ("base","ov2"):
    try:
        d=json.load(open(f"gpurun_out/r2y/bench_{nm}.json"))
    except Exception as e:
        print(nm, "unreadable", e); continue
    r=d["roofline"]
    print(nm, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
PY
