mkdir -p gpurun_out/r2o
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2o/bench_n${N}.json 2> gpurun_out/r2o/bench_n${N}.err; echo "bench rc=$?"; tail -2 gpurun_out/r2o/bench_n${N}.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2o/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["phase_ms_per_step"].items()}, (d.get("parity") or {}).get("vs_single_gpu"), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
