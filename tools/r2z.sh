O=gpurun_out/r2z; mkdir -p $O
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_dist.py -m gpu -x -q -k "ipc or nccl" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench rc=$?"; tail -3 $O/bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2z/bench_n2.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), d.get("parity"), {k:round(v,3) for k,v in r["phase_ms_per_step"].items()})
PY
