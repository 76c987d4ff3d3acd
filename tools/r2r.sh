mkdir -p gpurun_out/r2r
N=$(nvidia-smi -L | wc -l)
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/r2r/$name.json 2> gpurun_out/r2r/$name.err; echo "$name rc=$?"; grep -vE "OMP_NUM|^\*\*\*" gpurun_out/r2r/$name.err | tail -3; }
MXE_TIMING_FINE=1 run fine_n$N --steps 5 --warmup 3 --no-cpu-baseline
run c4_n$N --workload c4 --steps 3 --warmup 2 --no-cpu-baseline
run sweep_n$N --sweep --steps 3 --warmup 2 --no-cpu-baseline --no-parity
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2r/*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        if "sweep" in d:
            for p in d["sweep"]:
                print("sweep", p["k"], p["w"], round(p["value"],1), round(p["ms_per_step"],2), p["roofline_kernel"], round(p["roofline_frac"] or 0,3), round(p["sketch_frac"] or 0,3))
        else:
            print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["phase_ms_per_step"].items()}, (d.get("parity") or {}).get("vs_single_gpu"), d["config"]["minimizers"], d["config"]["vertices"], d["config"]["edges"])
    except Exception as e:
        print(f, "failed", e)
PY
