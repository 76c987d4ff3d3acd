mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_p2p.py -m gpu -x -q > gpurun_out/r2n/pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2n/pytest.log
N=$(nvidia-smi -L | wc -l)
if [ "$N" -gt 1 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n/bench_n${N}.json 2> gpurun_out/r2n/bench_n${N}.err; echo "bench rc=$?"; tail -2 gpurun_out/r2n/bench_n${N}.err
fi
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n/bench_n1.json 2> gpurun_out/r2n/bench_n1.err; echo "bench1 rc=$?"; tail -2 gpurun_out/r2n/bench_n1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, (d.get("parity") or {}).get("vs_single_gpu"))
    except Exception as e:
        print(f, "failed", e)
PY
