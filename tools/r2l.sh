mkdir -p gpurun_out/r2l
nvidia-smi topo -m > gpurun_out/r2l/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_p2p.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2l/pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2l/pytest.log
for mode in p2p allreduce; do
MXE_DIST_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l/bench_n2_$mode.json 2> gpurun_out/r2l/bench_n2_$mode.err; echo "bench $mode rc=$?"; tail -3 gpurun_out/r2l/bench_n2_$mode.err
done
python - <<'PY'
import json
for m in ("p2p","allreduce"):
    try:
        d=json.loads(open(f"gpurun_out/r2l/bench_n2_{m}.json").read().strip().split("\n")[-1])
        print(m, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()}, d.get("parity"))
    except Exception as e:
        print(m, "failed", e)
PY
