mkdir -p gpurun_out/r2m
N=$(nvidia-smi -L | wc -l)
for rec in 1 0; do
MXE_P2P_RECORDS=$rec timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2m/bench_n${N}_rec$rec.json 2> gpurun_out/r2m/bench_n${N}_rec$rec.err; echo "bench rec=$rec rc=$?"; tail -2 gpurun_out/r2m/bench_n${N}_rec$rec.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m/bench_n*_rec*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, (d.get("parity") or {}).get("vs_single_gpu"))
    except Exception as e:
        print(f, "failed", e)
PY
