O=gpurun_out/final; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench rc=$?"
python - <<'PY'
import json, glob, subprocess
n=int(subprocess.check_output("nvidia-smi -L | wc -l", shell=True))
d=json.loads(open(f"gpurun_out/final/bench_n{n}.json").read().strip().split("\n")[-1])
print(n, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["phase_ms_per_step"].items()}, (d.get("parity") or {}).get("vs_single_gpu"))
PY
