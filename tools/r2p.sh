mkdir -p gpurun_out/r2p
python bench.py --steps 5 --warmup 3 > gpurun_out/r2p/bench_n1.json 2> gpurun_out/r2p/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p/bench_ref.json 2> gpurun_out/r2p/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2p/bench_n1.json"))
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", d["e2e"])
print({k:round(v,2) for k,v in d["roofline"]["phase_ms_per_step"].items()})
r=d["roofline"]; print({k:r[k] for k in ("kernel","frac","pack_cand_frac","sketch_frac")}, r["alu"])
print("t3", json.dumps(d.get("t3"), indent=1))
print("cpu", d.get("cpu_baseline"))
print(open("gpurun_out/r2p/bench_ref.json").read()[:600])
PY
