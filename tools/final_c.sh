O=gpurun_out/final; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; grep -vE "OMP_NUM|^\*\*\*" $O/$name.err | tail -2; }
run bench_n$N --steps 5 --warmup 3 --no-cpu-baseline
run bench_n${N}_c4 --workload c4 --steps 3 --warmup 2 --no-cpu-baseline
run sweep_n$N --sweep --steps 2 --warmup 1 --no-cpu-baseline --no-parity
timeout 600 python -m pytest tests/test_gpu_p2p.py -m gpu -x -q -k "ipc" > $O/pytest_multigpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_multigpu.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/final/*_n8*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        if "sweep" in d:
            for p in d["sweep"]:
                print("sweep", p["k"], p["w"], round(p["value"],1), round(p["ms_per_step"],2), p["roofline_kernel"], round(p["roofline_frac"] or 0,3), round(p["pack_cand_frac"] or 0,3), round(p["sketch_frac"] or 0,3))
        else:
            print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["phase_ms_per_step"].items()}, d.get("parity"))
    except Exception as e:
        print(f, "failed", e)
PY
