mkdir -p gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_p2p.py -m gpu -x -q > gpurun_out/r2s/pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2s/pytest.log
N=$(nvidia-smi -L | wc -l)
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2s/$name.json 2> gpurun_out/r2s/$name.err; echo "$name rc=$?"; grep -vE "OMP_NUM|^\*\*\*" gpurun_out/r2s/$name.err | tail -3; }
if [ "$N" -gt 1 ]; then
MXE_TIMING_FINE=1 run fine_n$N --steps 5 --warmup 3 --no-cpu-baseline
run plain_n$N --steps 5 --warmup 3 --no-cpu-baseline
run c4_n$N --workload c4 --steps 3 --warmup 2 --no-cpu-baseline
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s/*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["phase_ms_per_step"].items()}, (d.get("parity") or {}).get("vs_single_gpu"))
    except Exception as e:
        print(f, "failed", e)
PY
