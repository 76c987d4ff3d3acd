#!/usr/bin/env python3
"""bench.py -- Gbases/s sketched+filtered at k=32 w=1000 (BASELINE.json metric).

A "step" = one pass of the hot path over the whole workload: ordered minimizer sketch of every
assembly (reference + target), per-assembly uniqueness, found-in-all intersection and the weighted
adjacent-minimizer edge list, with the result (flags, vertices, edges) delivered to host memory.

  value  : device-resident inputs (ASCII bases already in HBM), whole-job bases / step time
  e2e    : same step through the public host API (Engine.sketch_buffers: pinned host buffers,
           host->device copy inside the timed region, results read back)
  roofline / cpu_baseline : see DESIGN.md "Measurement"

N>1 (torchrun, one rank per GPU): records are sharded over ranks in contiguous ranges (strong
scaling, total work fixed).  Steps 2-3 run distributed (ntjoin_b200.dist): one NCCL all-gather of the
per-rank minimizer hashes (full multiset: uniqueness is per assembly, not per GPU), then every rank
marks its hash range and processes its own records, combined by three integer all-reduces; each rank
ends with its shard of the result (flags of its records, vertices of its hash range, its edges).

`--impl reference` times the CPU restatement of the reference path (oracle/, all host threads) on
a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K, W = 32, 1000               # overridden per run from the workload / -k / -w
WEIGHTS = [2.0, 1.0]          # reference_weights='2', target_weight=1 (tests/ntjoin_test.py:22); one 2.0 per reference
GRCH38_MBP = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4"],
                    help="c3: BASELINE configs[2] 3 Gbp target + 3 Gbp reference (default); c2: configs[1] 100 Mbp + 100 Mbp; "
                         "c4: configs[3] 3 Gbp target + 3 references, w=500 (multi-way intersection)")
    ap.add_argument("-k", type=int, default=0, help="k-mer size (default: the workload's, 32)")
    ap.add_argument("-w", type=int, default=0, help="window size (default: the workload's)")
    ap.add_argument("--sweep", action="store_true",
                    help="configs[4]: k in {24,32,40} x w in {250,500,1000,5000} on the workload's data, device-resident timing and "
                         "roofline fraction per point (one JSON line with a 'sweep' list; single GPU)")
    ap.add_argument("--bases", type=float, default=0, help="override bases per assembly")
    ap.add_argument("--with-n", action="store_true", help="put 0.5%% of the reference in N runs (default: N-free headline variant)")
    ap.add_argument("--cpu-sample", type=float, default=0, help="bases per assembly for the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-t3", action="store_true", help="skip the file-seam measurement (FASTA file -> TSV file -> Python objects)")
    ap.add_argument("--t3-bases", type=float, default=200e6, help="bases per assembly of the file-seam measurement")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the comparison of the merged shards with the single-GPU result")
    return ap.parse_args()


def workload_spec(args):
    n_refs, w = 1, 1000
    if args.workload == "c2":
        g, nchr, prop, lo, hi, dup = 100e6, 10, None, 20_000, 2_000_000, 0.0
        name = "configs[1]: synthetic 100 Mbp target + 100 Mbp reference"
    elif args.workload == "c4":
        g, nchr, prop, lo, hi, dup = 3e9, 24, GRCH38_MBP, 50_000, 20_000_000, 0.02
        n_refs, w = 3, 500
        name = "configs[3]: synthetic 3 Gbp target + 3 references (multi-way intersection)"
    else:
        g, nchr, prop, lo, hi, dup = 3e9, 24, GRCH38_MBP, 50_000, 20_000_000, 0.02
        name = "configs[2]: synthetic 3 Gbp human-scale target + 1 reference"
    if args.bases:
        g = args.bases
        name += f" (scaled to {g:.3g} bp per assembly)"
    return dict(G=int(g), n_chrom=nchr, prop=prop, lo=lo, hi=hi, dup=dup, name=name, n_refs=n_refs, w=w)


# ------------------------------------------------------------------------------------ data (GPU)
def gen_assemblies_gpu(spec, with_n, device):
    """Reference (i.i.d. ACGT, duplicated 5 kb segments) and a target derived from it (cut into contigs
    with log-uniform lengths, half reverse-complemented, 0.1 % substitutions, shuffled).  Generated on
    the device with a fixed seed; identical on every rank."""
    import torch
    G = spec["G"]
    g = torch.Generator(device=device)
    g.manual_seed(20251017)
    rng = np.random.Generator(np.random.PCG64(20251017))
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    ref = torch.empty(G, dtype=torch.uint8, device=device)
    step = 1 << 28
    for s in range(0, G, step):
        e = min(G, s + step)
        ref[s:e] = lut[torch.randint(0, 4, (e - s,), dtype=torch.uint8, device=device, generator=g).long()]
    prop = np.asarray(spec["prop"] or [1.0] * spec["n_chrom"], dtype=np.float64)[:spec["n_chrom"]]
    lens = np.maximum(1, (prop / prop.sum() * G).astype(np.int64))
    lens[-1] += G - lens.sum()
    roffs = np.zeros(len(lens) + 1, dtype=np.uint64)
    roffs[1:] = np.cumsum(lens)
    if spec["dup"] > 0:
        seg = 5000
        for _ in range(max(1, int(G * spec["dup"] / seg / 6))):
            s = int(rng.integers(0, G - seg))
            for _c in range(int(rng.integers(1, 10))):
                d = int(rng.integers(0, G - seg))
                ref[d:d + seg] = ref[s:s + seg].clone()
    if with_n:
        done, target = 0, int(G * 0.005)
        while done < target:
            ln = int(min(target - done + 100, np.exp(rng.uniform(np.log(100), np.log(50000)))))
            s = int(rng.integers(0, G - ln))
            ref[s:s + ln] = ord("N")
            done += ln
    pieces = []
    for c in range(len(lens)):
        at, e = int(roffs[c]), int(roffs[c + 1])
        while at < e:
            ln = min(int(np.exp(rng.uniform(np.log(spec["lo"]), np.log(spec["hi"])))), e - at)
            pieces.append((at, at + ln))
            at += ln
    order = rng.permutation(len(pieces))
    tgt = torch.empty(G, dtype=torch.uint8, device=device)
    toffs = np.zeros(len(pieces) + 1, dtype=np.uint64)
    at = 0
    for i, pi in enumerate(order):
        a, b = pieces[pi]
        chunk = ref[a:b]
        if rng.random() < 0.5:
            chunk = comp[chunk.flip(0).long()]
        tgt[at:at + (b - a)] = chunk
        at += b - a
        toffs[i + 1] = at
    n_sub = int(G * 0.001)
    idx = torch.unique(torch.randint(0, G, (n_sub,), device=device, generator=g))   # unique: deterministic scatter
    sub = lut[torch.randint(0, 4, (idx.numel(),), dtype=torch.uint8, device=device, generator=g).long()]
    keep = tgt[idx] != ord("N")
    tgt[idx[keep]] = sub[keep]
    refs = [(ref, roffs)]
    for r in range(1, spec.get("n_refs", 1)):       # further references: the ancestor with 0.1-0.5 % independent substitutions
        other = ref.clone()
        n_s = int(G * 0.001 * (r + 1))
        ix = torch.unique(torch.randint(0, G, (n_s,), device=device, generator=g))
        sb = lut[torch.randint(0, 4, (ix.numel(),), dtype=torch.uint8, device=device, generator=g).long()]
        ok = other[ix] != ord("N")
        other[ix[ok]] = sb[ok]
        refs.append((other, roffs))
    return refs + [(tgt, toffs)]     # assembly order: reference(s) first, target last


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU every few ms DURING the timed region
    (NVML in a thread; falls back to `nvidia-smi -lms` when pynvml is unavailable)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.proc, self.t, self.nvml = index, [], False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), float(mx), pw, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 7 and r[0].replace(".", "").isdigit():
                self.rows.append((float(r[0]), float(r[1]), float(r[2]) if r[2].replace(".", "").isdigit() else 0.0,
                                  [names[i] for i in range(4) if r[3 + i].lower() == "active"]))

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": max(r[1] for r in self.rows),
                "power_w_max": max(r[2] for r in self.rows), "samples": len(self.rows),
                "reasons": sorted({x for r in self.rows for x in r[3]})}


def find_real_indexlr():
    """a btllib `indexlr` that is neither this repo's drop-in nor the oracle's CLI (prefer it as the CPU
    baseline and byte-compare against it when one exists; tests/test_real_indexlr.py does the comparison)"""
    import glob
    import shutil
    ours = {os.path.realpath(os.path.join(ROOT, "bin", "indexlr")), os.path.realpath(os.path.join(ROOT, "oracle", "_build", "mxo_indexlr"))}
    for c in [shutil.which("indexlr")] + glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", "indexlr"), recursive=True):
        if c and os.path.isfile(c) and os.access(c, os.X_OK) and os.path.realpath(c) not in ours:
            try:
                head = open(c, "rb").read(4096)
            except OSError:
                continue
            if b"ntjoin_b200" in head or b"mxo_indexlr" in head:
                continue
            return c
    return None


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(spec, args, steps, warmup, sample_bases=None, with_t4=False):
    """The reference path on host cores: oracle sketch (one record per worker thread, like indexlr -t)
    + oracle steps 2-3, on a bounded sample of the same workload (same generator, smaller G).
    with_t4: also time the sketch with 4 threads, the reference's default (`t=4`, ntJoin:48)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from ntjoin_b200 import synth
    orc = oracle_lib.Oracle()
    cores = os.cpu_count() or 1
    if sample_bases is None:
        sample_bases = args.cpu_sample or min(spec["G"], 25e6 * cores)   # ~1 s per core-step at ~30 Mbases/s/thread
    sample_bases = int(sample_bases)
    n_chrom = max(spec["n_chrom"], cores)         # enough records to keep every thread busy
    rseq, roffs, _ = synth.make_reference(sample_bases, n_chrom=n_chrom, dup_frac=spec["dup"])
    tseq, toffs, _ = synth.derive_target(rseq, roffs, min_len=spec["lo"], max_len=min(spec["hi"], max(spec["lo"] * 2, sample_bases // cores)))
    asms = [(rseq, roffs)] * spec.get("n_refs", 1) + [(tseq, toffs)]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        sks = [orc.sketch(s, o, K, W, threads=cores) for s, o in asms]
        orc.filter_and_edges([m["out_hash"] for m in sks], [m["contig"] for m in sks], WEIGHTS)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = len(asms) * sample_bases
    sec = float(np.mean(times))
    real = find_real_indexlr()
    out = {"value": total / sec / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
           "real_indexlr": real or "none found (probed: PATH, baseline/_ref/); the oracle restatement stands in",
           "sample": f"{sample_bases} bp reference + derived target (same generator as the workload), k={K} w={W}, "
                     f"{steps} step(s) after {warmup} warm-up, oracle/mxo.c with {cores} threads"}
    if with_t4:
        # the reference's default thread count, on a quarter of the sample (bounded CPU time)
        sub = [(s[:int(o[max(1, len(o) // 4)])], o[:max(1, len(o) // 4) + 1]) for s, o in asms]
        t0 = time.perf_counter()
        sk4 = [orc.sketch(s, o, K, W, threads=4) for s, o in sub]
        orc.filter_and_edges([m["out_hash"] for m in sk4], [m["contig"] for m in sk4], WEIGHTS)
        dt = time.perf_counter() - t0
        out["t4"] = {"value": sum(int(o[-1]) for _, o in sub) / dt / 1e9, "unit": "Gbases/s", "cores": 4,
                     "note": "oracle sketch with 4 threads (the reference's default t=4, ntJoin:48) + oracle steps 2-3, first quarter of the sample's records"}
    return out, sec


def run_reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = max(1, min(args.warmup, 1))
    steps = max(1, min(args.steps, 3))
    cb, sec = cpu_reference_run(spec, args, steps, warm)
    line = {"metric": f"Gbases/s sketched+filtered at k={K} w={W}", "value": cb["value"], "unit": "Gbases/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": spec["name"], "k": K, "w": W, "note": "bounded sample on host cores"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ the bound that actually binds
ALU_OPS_PER_BASE = {"cand31_kernel": 9.1,      # ALU-pipe instructions per base, counted in the SASS
                    "scan_bs2_kernel": 3.9}    # ALU-pipe instructions of the 16-step loop body / 512 positions (tests/test_sass_counts.py);
                                               # the k-1 halo of every stream adds (k - 1) / L on top


def alu_roofline(bases_per_launch, launch_ms, kernel="scan_bs2_kernel"):
    """the candidate kernel against the INT32 ALU pipe (LOP3/SHF/IADD3/ISETP/PRMT): the secondary roofline of SURVEY.md 8(d).
    Peak = LOP3 rate measured on this GPU type by tools/alu_peak.cu (profiles/r01_l_alu_peak_microbench.jsonl), else the
    nominal 148 SMs x 64 lanes x 1.965 GHz."""
    try:
        peak, source = 148 * 64 * 1.965e9 / 1e12, "nominal 148 SM x 64 lanes x 1.965 GHz"
        try:
            with open(os.path.join(ROOT, "profiles", "r01_l_alu_peak_microbench.jsonl")) as fh:
                for row in fh:
                    r = json.loads(row)
                    if str(r.get("op", "")).startswith("LOP3"):
                        peak, source = float(r["tera_thread_ops_per_s"]), "tools/alu_peak.cu LOP3 rate (profiles/r01_l_alu_peak_microbench.jsonl)"
        except (OSError, ValueError, KeyError):
            pass
        opb = ALU_OPS_PER_BASE.get(kernel, 9.1)
        achieved = opb * bases_per_launch / (launch_ms * 1e-3) / 1e12 if launch_ms > 0 else 0.0
        return {"bound": "int32 alu pipe", "kernel": kernel, "ops_per_base": opb, "achieved": achieved, "peak": peak, "unit": "Tops/s",
                "frac": achieved / peak if peak else None, "peak_source": source}
    except Exception as exc:           # reporting only: never lose the bench line over it
        return {"error": str(exc)}


# ------------------------------------------------------------------------------------ k x w sweep (configs[4])
def run_sweep(args, spec, eng, world, rank, step_device, timed, stats, my_bases, total_bases, n_asm, dist_mode):
    """configs[4]: k in {24, 32, 40} x w in {250, 500, 1000, 5000}.  Per point the device-resident step (whole job, all
    ranks; max over ranks like the headline number) and the roofline fractions of the sketch at its three scopes."""
    global K, W
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    points = []
    for k in (24, 32, 40):
        for w in (250, 500, 1000, 5000):
            K, W = k, w
            for _ in range(max(1, min(args.warmup, 2))):
                step_device()
            eng.timing_reset()
            sec = timed(step_device, args.steps) / args.steps
            if rank != 0:
                continue
            algo = (my_bases + 16.0 * sum(stats["n_mx"])) / n_asm          # algorithmic bytes per assembly of this rank
            kt = {}
            for name, span in (("pack2_kernel", "k_pack2"), ("scan_bs2_kernel", "k_scan")):
                t_k, n_k = eng.timing(span)
                if n_k:
                    kt[name] = t_k / n_k
            t_sketch, n_sk = eng.timing("sketch")
            t_pack, t_cand = eng.timing("pack")[0], eng.timing("cand")[0]
            asm_per_call = n_asm if n_sk and n_sk <= args.steps * 1.5 else 1       # one sketch call may cover all assemblies
            per_asm = lambda t: t / args.steps / n_asm                              # noqa: E731
            dom = max(kt, key=kt.get) if kt else None
            points.append({"k": k, "w": w, "value": total_bases / sec / 1e9, "ms_per_step": sec * 1e3,
                           "kernel_ms": {n: v / asm_per_call for n, v in kt.items()},
                           "roofline_kernel": dom, "roofline_frac": algo * asm_per_call / (kt[dom] * 1e-3) / 1e9 / peak if dom else None,
                           "pack_cand_frac": algo / (per_asm(t_pack + t_cand) * 1e-3) / 1e9 / peak if t_pack + t_cand > 0 else None,
                           "sketch_frac": algo / (per_asm(t_sketch) * 1e-3) / 1e9 / peak if t_sketch > 0 else None,
                           "minimizers_rank0": stats["n_mx"], "vertices_rank0": stats["vertices"], "edges_rank0": stats["edges"],
                           "sketch_ms": t_sketch / args.steps, "filter_ms": eng.timing("filter")[0] / args.steps})
    if rank == 0:
        print(json.dumps({"metric": "Gbases/s sketched+filtered, k x w sweep", "unit": "Gbases/s", "n_gpus": world, "steps": args.steps,
                          "config": {"workload": spec["name"], "bases_per_step": total_bases, "dist_mode": dist_mode if world > 1 else None},
                          "roofline_peak_gbs": peak, "sweep": points}), flush=True)


# ------------------------------------------------------------------------------------ T3: the file seam
def run_t3(args, eng, spec, dev):
    """what a user of `bin/with-b200` gets.  Page-cache-warm FASTA files -> `<fasta>.k.w.tsv` on disk
    (Engine.sketch_file + write_tsv = bin/indexlr, seams S1/S2) -> the drop-in read_minimizers / filter_minimizers /
    build_graph (seam S3: Python dict / list / graph objects exactly as bin/ntjoin_utils.py returns them).  Beside it,
    on the same files and host: the CPU sketch (oracle/mxo_indexlr.c; the reference's own is an un-vendored binary) with
    the reference's default 4 threads (ntJoin:48) and with all cores, then steps 2-3 in plain Python as the reference
    does them (tests/ref_py.py follows bin/ntjoin_utils.py:167-193, :152-165, :83-141 statement by statement; the
    reference itself is not on the GPU box).  A bounded workload (--t3-bases per assembly): the Python side costs
    ~3 us per minimizer."""
    import shutil
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import ref_py
    from ntjoin_b200 import dropin
    oracle_lib.build()
    small = dict(spec)
    small["G"] = int(args.t3_bases)
    small["hi"] = min(spec["hi"], max(spec["lo"] * 2, small["G"] // 16))
    asms = gen_assemblies_gpu(small, False, dev)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="mxe_t3_", dir=base)
    out = {"workload": f"{len(asms)} assemblies x {small['G']} bp (same generator as the headline workload), k={K} w={W}, FASTA in {base or 'tmp'} (page-cache warm)"}
    try:
        paths = []
        for i, (seq, offs) in enumerate(asms):
            host = seq.cpu().numpy()
            path = os.path.join(tmp, ("ref%d.fa" % i) if i + 1 < len(asms) else "target.fa")
            with open(path, "wb") as fh:
                for c in range(len(offs) - 1):
                    fh.write(b">ctg%06d\n" % c)
                    fh.write(host[int(offs[c]):int(offs[c + 1])].tobytes())
                    fh.write(b"\n")
            paths.append(path)
            del host
        del asms
        total = small["G"] * len(paths)
        weights = {p + f".k{K}.w{W}.tsv": w for p, w in zip(paths, WEIGHTS)}
        mod = dropin.install(ref_py.as_module())
        dropin._ENGINE = eng                                   # the bench's engine serves the drop-in layer

        def engine_run():
            t = {}
            t0 = time.perf_counter()
            for p in paths:
                sk = eng.sketch_file(p, K, W)
                sk.write_tsv(p + f".k{K}.w{W}.tsv", pos=True, strand=False, seq=True)
                sk.close()
            t["fasta_to_tsv_s"] = time.perf_counter() - t0
            t1 = time.perf_counter()
            list_mx_info, list_mxs = {}, {}
            for p in paths:
                tsv = p + f".k{K}.w{W}.tsv"
                list_mx_info[tsv], list_mxs[tsv] = mod.read_minimizers(tsv)
            list_mxs = mod.filter_minimizers(list_mxs)
            g = mod.build_graph(list_mxs, weights)
            t["tsv_to_graph_objects_s"] = time.perf_counter() - t1
            t["total_s"] = time.perf_counter() - t0
            t["edges"] = len(g.es)
            return t

        engine_run()                                            # warm-up (page cache, allocations)
        t = engine_run()
        out["engine"] = dict(t, gbases_per_s=total / t["total_s"] / 1e9)
        cores = os.cpu_count() or 1
        baseline = {}
        for label, threads in (("t4", 4), ("all", cores)):
            t0 = time.perf_counter()
            for p in paths:
                subprocess.check_call([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", str(K), "-w", str(W), "-t", str(threads),
                                       "-o", p + ".cpu.tsv", p])
            baseline[f"sketch_{label}_s"] = time.perf_counter() - t0
            baseline[f"sketch_{label}_threads"] = threads
        orig = ref_py.as_module()
        t0 = time.perf_counter()
        lm = {}
        for p in paths:
            _info, lm[p + ".cpu.tsv"] = orig.read_minimizers(p + ".cpu.tsv")
        lm = orig.filter_minimizers(lm)
        gb = orig.build_graph(lm, {p + ".cpu.tsv": w for p, w in zip(paths, WEIGHTS)})
        baseline["python_steps23_s"] = time.perf_counter() - t0
        baseline["edges"] = len(gb.es)
        baseline["total_t4_s"] = baseline["sketch_t4_s"] + baseline["python_steps23_s"]
        baseline["total_all_s"] = baseline["sketch_all_s"] + baseline["python_steps23_s"]
        baseline["gbases_per_s_t4"] = total / baseline["total_t4_s"] / 1e9
        baseline["kind"] = "port (oracle/mxo_indexlr.c + tests/ref_py.py): the reference's own step 1 is an un-vendored binary, its bin/ is not on this box"
        out["cpu"] = baseline
        out["graph_note"] = ("python-igraph is not installable here: both arms hand their vertices and edges to tests/ref_py.RecordingGraph, which "
                             "only stores what it is given (as the C library would) and resolves names / edge ids on demand")
        out["same_edge_count"] = baseline["edges"] == t["edges"]
        out["speedup_vs_t4"] = baseline["total_t4_s"] / t["total_s"]
        out["speedup_vs_all_cores"] = baseline["total_all_s"] / t["total_s"]
    finally:
        dropin._ENGINE = None
        dropin._LAST_FILTER = None
        shutil.rmtree(tmp, ignore_errors=True)
    return out


# ------------------------------------------------------------------------------------ N > 1 vs N = 1
def result_digest(d):
    """order-sensitive digest of a whole result (minimizer tuples, flags, vertices, edge list)"""
    import hashlib
    h = hashlib.sha256()
    for key in ("out_hash", "pos", "contig", "forward", "uniq", "keep"):
        for a in d[key]:
            h.update(np.ascontiguousarray(a).astype(np.uint64 if key == "out_hash" else np.uint32 if key in ("pos", "contig") else np.uint8).tobytes())
    for key, dt in (("vertices", np.uint64), ("edge_u", np.uint64), ("edge_v", np.uint64), ("support", np.uint32), ("weight", np.float64)):
        h.update(np.ascontiguousarray(d[key], dtype=dt).tobytes())
    return h.hexdigest()[:16]


def multi_gpu_parity(eng, shards, gather_and_filter, single, rank, world, dist, dist_mode):
    """One more step outside the timed regions: every rank contributes its minimizers and its fetched result shard,
    rank 0 merges them (ntjoin_b200.dist.merge_shards) and compares field by field with the single-GPU result of the
    same job.  A mismatch fails the run."""
    from ntjoin_b200.dist import merge_shards
    sks = [eng.sketch_device(s.data_ptr(), o, K, W) for s, o, _ in shards]
    res = gather_and_filter(sks)
    mine = {"sk": [(sk.out_hash.copy(), sk.pos.copy(), sk.contig.copy() + np.uint32(c0), sk.forward.copy()) for sk, (_, _, c0) in zip(sks, shards)],
            "shard": {k: ([x.copy() for x in v] if isinstance(v, list) else np.array(v, copy=True)) for k, v in res.fetch().items()}}
    for sk in sks:
        sk.close()
    res.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    n_asm = len(shards)
    merged = merge_shards([p["shard"] for p in parts])
    merged["out_hash"] = [np.concatenate([p["sk"][a][0] for p in parts]) for a in range(n_asm)]
    merged["pos"] = [np.concatenate([p["sk"][a][1] for p in parts]) for a in range(n_asm)]
    merged["contig"] = [np.concatenate([p["sk"][a][2] for p in parts]) for a in range(n_asm)]
    merged["forward"] = [np.concatenate([p["sk"][a][3] for p in parts]) for a in range(n_asm)]
    bad = []
    for key in ("out_hash", "pos", "contig", "forward", "uniq", "keep"):
        for a in range(n_asm):
            if not np.array_equal(np.asarray(merged[key][a]).astype(np.uint64), np.asarray(single[key][a]).astype(np.uint64)):
                bad.append(f"{key}[{a}]")
    for key in ("vertices", "edge_u", "edge_v", "support", "weight"):
        if not np.array_equal(merged[key], single[key]):
            bad.append(key)
    return {"vs_single_gpu": not bad, "mismatch": bad, "digest": result_digest(merged), "single_gpu_digest": result_digest(single),
            "dist_mode": dist_mode, "compared": "all minimizer tuples, uniq/keep flags, vertices, edge list (order, support, weight)"}


# ------------------------------------------------------------------------------------ GPU arm
def main():
    global K, W, WEIGHTS
    args = parse_args()
    spec = workload_spec(args)
    K, W = args.k or 32, args.w or spec["w"]
    WEIGHTS = [2.0] * spec["n_refs"] + [1.0]
    if args.impl == "reference":
        run_reference_arm(args, spec)
        return

    import torch
    import torch.distributed as dist
    import ntjoin_b200
    from ntjoin_b200.dist import DeviceArray, TorchComm, distributed_filter_and_edges, shard_ranges

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    eng = ntjoin_b200.Engine(local, timing=True)
    fine = os.environ.get("MXE_TIMING_FINE", "0") not in ("", "0")       # per-kernel times of steps 2-3 (costs two event records per launch)
    if fine:
        eng.set_option("timing", 2)
    stream = torch.cuda.Stream()          # one real stream for the engine, the torch ops and the timing events
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    comm = TorchComm(dev, timing=True) if world > 1 else None
    # steps 2-3 across ranks: "p2p" = direct peer stores over NVLink + device-side barriers (csrc/p2p.cu, default);
    # "allreduce" / "alltoall" = the NCCL formulations of round 1 (kept for comparison)
    dist_mode = os.environ.get("MXE_DIST_MODE", "p2p")
    stages = (eng.a2a_stages() if dist_mode == "alltoall" else eng.dist_stages()) if world > 1 and dist_mode != "p2p" else None
    p2p = None
    if world > 1 and dist_mode == "p2p":
        cap_total = int((len(WEIGHTS)) * spec["G"] * 2.0 / ((250 if args.sweep else W) + 1) * 1.15) + 100_000
        p2p = eng.p2p(rank, world, cap_total, n_asm_max=max(4, len(WEIGHTS)))
        handles = [None] * world
        dist.all_gather_object(handles, p2p.handle())          # 64 bytes per rank, once
        p2p.connect(handles)

    assemblies = gen_assemblies_gpu(spec, args.with_n, dev)
    my = shard_ranges([o for _, o in assemblies], world)[rank]
    shards = []            # per assembly: (device tensor, local offsets, first global record id)
    for (seq, offs), (c0, c1) in zip(assemblies, my):
        lo, hi = int(offs[c0]), int(offs[c1])
        local_seq = seq[lo:hi].clone() if world > 1 else seq
        shards.append((local_seq, (offs[c0:c1 + 1] - offs[c0]).astype(np.uint64), c0))
    parity = None
    single = None
    if world > 1 and rank == 0 and not args.no_parity:
        # the whole job once on ONE GPU (rank 0 holds all assemblies at this point): the result every N must reproduce
        sks = [eng.sketch_device(seq.data_ptr(), offs, K, W) for seq, offs in assemblies]
        res = eng.filter_and_edges(sks, WEIGHTS)
        f = res.fetch()
        single = {"out_hash": [sk.out_hash.copy() for sk in sks], "pos": [sk.pos.copy() for sk in sks],
                  "contig": [sk.contig.copy() for sk in sks], "forward": [sk.forward.copy() for sk in sks],
                  "uniq": [u.copy() for u in f["uniq"]], "keep": [k.copy() for k in f["keep"]], "vertices": f["vertices"].copy(),
                  "edge_u": f["edge_u"].copy(), "edge_v": f["edge_v"].copy(), "support": f["support"].copy(), "weight": f["weight"].copy()}
        for sk in sks:
            sk.close()
        res.close()
    del assemblies
    if world > 1:
        torch.cuda.empty_cache()
    # all assemblies of this rank in ONE device buffer (every assembly 16-byte aligned): the device-resident step sketches
    # them in one call (Engine.sketch_device_multi)
    starts, at = [], 0
    for s_, _, _ in shards:
        starts.append(at)
        at += (s_.numel() + 15) // 16 * 16
    combo = torch.full((max(16, at),), ord("N"), dtype=torch.uint8, device=dev)
    for i, (s_, o_, c_) in enumerate(shards):
        combo[starts[i]:starts[i] + s_.numel()] = s_
        shards[i] = (combo[starts[i]:starts[i] + s_.numel()], o_, c_)
    del s_
    if world > 1:
        torch.cuda.empty_cache()
    my_bases = sum(int(o[-1]) for _, o, _ in shards)
    total_bases = spec["G"] * len(shards)
    host = [torch.empty(s.numel(), dtype=torch.uint8).pin_memory() for s, _, _ in shards]
    for h, (s, _, _) in zip(host, shards):
        h.copy_(s)
    torch.cuda.synchronize()

    n_asm = len(shards)
    stats = {}

    def gather_and_filter(sks):
        if world == 1:
            res = eng.filter_and_edges(sks, WEIGHTS)
            return res
        if p2p is not None:
            from ntjoin_b200.dist import DistShard
            return DistShard(p2p.run(sks, WEIGHTS), None, rank, None, None)      # this rank's shard of the result
        hashes, contigs = [], []
        for sk in sks:
            n, ph, _pp, pc = sk.device_pointers()
            hashes.append(torch.as_tensor(DeviceArray(ph, n, "<i8"), device=dev) if n else torch.empty(0, dtype=torch.int64, device=dev))
            contigs.append(torch.as_tensor(DeviceArray(pc, n, "<i4"), device=dev) if n else torch.empty(0, dtype=torch.int32, device=dev))
        return distributed_filter_and_edges(stages, hashes, contigs, WEIGHTS, comm)   # this rank's shard of the result

    # one sketch call for all assemblies of the rank while their valid k-mer ordinals fit 32 bits (the window selection
    # has a 32-bit variant: measured 1.13 vs 1.57 ms per step on 6 Gbp), else one call per assembly
    n_records = sum(len(o) for _, o, _ in shards)

    # Large shares (one sketch per assembly): all assemblies are enqueued back to back on the engine stream and their counts
    # are read in ONE host round trip at the end (Engine.sketch_device_many, option many_streams = 1), so the GPU does not
    # idle while the host reads the sizes of assembly i and enqueues assembly i+1: 10.17 vs 10.22 ms per step on configs[2].
    # MXE_SKETCH_OVERLAP=1 alternates the assemblies between two streams instead (measured 12.04 vs 11.90 ms in round 2a:
    # every kernel of the sketch already fills the machine with CTAs, the second stream only gets SMs when the first
    # drains); MXE_SKETCH_OVERLAP=0 = one blocking call per assembly.
    overlap_mode = os.environ.get("MXE_SKETCH_OVERLAP", "2")
    overlap = overlap_mode not in ("", "0")
    eng.set_option("many_streams", 2 if overlap_mode == "1" else 1)

    def step_device(concurrent=True):
        # small per-rank shares: ONE sketch call over all assemblies (fewer launches and round trips; needs the valid k-mer
        # ordinals to fit 32 bits).  Large ones: one sketch per assembly, all enqueued before the sizes are read
        # (Engine.sketch_device_many); concurrent=False runs one blocking call per assembly (clean per-kernel times).
        use_multi = my_bases + (n_records + 8) * W < 0xF0000000
        if use_multi:
            parent, sks = eng.sketch_device_multi(combo.data_ptr(), [o for _, o, _ in shards], K, W, starts=starts)
        elif concurrent and overlap:
            parent, sks = None, eng.sketch_device_many([s.data_ptr() for s, _, _ in shards], [o for _, o, _ in shards], K, W)
        else:
            parent, sks = None, [eng.sketch_device(s.data_ptr(), o, K, W) for s, o, _ in shards]
        res = gather_and_filter(sks)
        n_tot, n_v, n_e = res.counts()        # result stays resident in HBM; sizes only
        stats["n_mx"] = [sk.n for sk in sks]
        stats["edges"], stats["vertices"] = n_e, n_v
        # e2e reads back: the minimizer tuples (out_hash u64, min_hash u64, pos u32, record u32, strand u8), the two flag
        # bytes per minimizer, the vertices and the weighted edge list (u, v, support mask, weight [+ order key on shards])
        stats["d2h"] = n_tot * 25 + n_tot * 2 + n_v * 8 + n_e * (28 if world == 1 else 36)
        for sk in ([parent] if parent is not None else sks):
            sk.close()
        res.close()

    def step_e2e():
        # H2D of assembly i+1 overlaps sketch i; the tuples of assembly i start their way back as soon as sketch i is done
        # (the other PCIe direction), beside the H2D / sketch of assembly i+1 and steps 2-3
        sks = eng.sketch_many([(h, o) for h, (_, o, _) in zip(host, shards)], K, W, prefetch_host=True)
        res = gather_and_filter(sks)
        for sk in sks:
            sk.fetch(copy=False)              # (out_hash, min_hash, pos, record, strand) of every minimizer into pinned host memory
        res.fetch(copy=False)                 # flags + vertices + weighted edge list into (pinned) host memory
        for sk in sks:
            sk.close()
        res.close()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        t = torch.tensor([max(wall, dev_ms / 1e3)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.sweep:
        run_sweep(args, spec, eng, world, rank, step_device, timed, stats, my_bases, total_bases, n_asm, dist_mode)
        if p2p is not None:
            dist.barrier()
            p2p.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        eng.close()
        return

    for _ in range(args.warmup):
        step_device()
    eng.timing_reset()
    if comm:
        comm.timing_ms()
    l0 = eng.kernel_launches()
    sampler = ClockSampler(local)
    sampler.start()
    sec = timed(step_device, args.steps)
    clocks = sampler.stop()
    launches = eng.kernel_launches() - l0
    t_filter_timed = eng.timing("filter")[0]
    # per-phase / per-kernel device times from a separate pass in which the assemblies are sketched one after the other
    # (in the timed region above their kernels overlap on two streams, so phase times there do not add up)
    eng.timing_reset()
    for _ in range(args.steps):
        step_device(concurrent=False)
    t_cand, n_cand = eng.timing("cand")
    t_pack, n_pack = eng.timing("pack")
    t_sketch, _ = eng.timing("sketch")
    t_filter, _ = eng.timing("filter")
    phases = {nm: eng.timing(nm)[0] / args.steps for nm in ("pack", "rank", "cand", "eval", "select", "gap", "emit", "sketch", "filter")}
    kernel_times = {kname: eng.timing(span) for kname, span in (("pack2_kernel", "k_pack2"), ("scan_bs2_kernel", "k_scan"))}
    if world == 1 or p2p is not None:
        phases.update({nm: eng.timing(nm)[0] / args.steps for nm in ("p2p_scatter", "p2p_buckets", "p2p_adjacency", "p2p_edges", "p2p_finish")})
        if fine:                 # every kernel of steps 2-3, and the waits at the device-side barriers (sketch-time skew lands in wait0)
            for nm in ("p2p_wait0", "p2p_wait1", "p2p_wait2", "p2p_wait3", "p2p_wait4", "k_p2p_scatter_kernel", "k_p2p_push_counts_kernel",
                       "k_p2p_bucket_kernel", "k_p2p_vbase_kernel", "k_p2p_flags_kernel", "k_p2p_compact_kernel", "k_p2p_succ_kernel",
                       "k_p2p_sight_kernel", "k_p2p_edge_owner_kernel", "k_p2p_table_kernel", "k_p2p_rec_owner_kernel",
                       "k_p2p_vertices_kernel", "k_p2p_first_count_kernel", "k_p2p_first_start_kernel", "k_p2p_edge_emit_kernel", "k_p2p_rec_emit_kernel"):
                t_k = eng.timing(nm)[0]
                if t_k:
                    phases[nm] = t_k / args.steps
    elif comm:
        names = ("a2a_partition", "a2a_mark", "a2a_sightings", "a2a_finish") if dist_mode == "alltoall" else \
            ("dist_mark", "dist_adjacency", "dist_edges", "dist_finish")
        phases.update({nm: eng.timing(nm)[0] / args.steps for nm in names})
        phases.update({"comm_" + k: v / args.steps for k, v in comm.timing_ms().items()})

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    sec_e2e = timed(step_e2e, args.steps)

    if world > 1 and not args.no_parity:
        parity = multi_gpu_parity(eng, shards, gather_and_filter, single, rank, world, dist, dist_mode)
    if world > 1:   # whole-job totals for the report (outside the timed regions)
        tot = torch.tensor(stats["n_mx"] + [stats["vertices"], stats["edges"], stats["d2h"], my_bases], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        tot = tot.cpu().tolist()
        stats["n_mx_total"], stats["vertices"], stats["edges"] = tot[:n_asm], tot[n_asm], tot[n_asm + 1]
        stats["d2h_total"], stats["h2d_total"] = tot[n_asm + 2], tot[n_asm + 3]

    parity_failed = False
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic bytes of the sketch: 1 B per base read + 16 B per emitted minimizer, whatever the
        # internal representation.  Three scopes, all on CUDA events of the engine stream over the timed region:
        #   per kernel   : one launch of the two front-end kernels (pack2_kernel reads the bases, scan_bs2_kernel decides
        #                  which positions can be minimizers); `frac` is the DOMINANT (longest) one of the two
        #   pack_cand    : both together (the packing pass inside the timed region)
        #   sketch       : whole step 1 per assembly (T1 of SURVEY 8(d))
        n_mx_local = sum(stats["n_mx"])
        algo_bytes_per_launch = (my_bases + 16.0 * n_mx_local) / n_asm
        traffic_tab = {}
        try:
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        except (OSError, ValueError):
            pass
        kernels = {}
        for kname, (t_k, n_k) in kernel_times.items():
            if n_k:
                ms = t_k / n_k
                tj = traffic_tab.get(kname)
                kernels[kname] = {"launch_ms": ms, "launches": int(n_k), "achieved": algo_bytes_per_launch / (ms * 1e-3) / 1e9,
                                  "frac": algo_bytes_per_launch / (ms * 1e-3) / 1e9 / peak,
                                  "traffic": (tj["dram_read_bytes_per_base"] + tj["dram_write_bytes_per_base"]) * my_bases / n_asm if tj else None}
        if kernels:
            dom = max(kernels, key=lambda kn: kernels[kn]["launch_ms"])
            cand_ms, achieved, traffic = kernels[dom]["launch_ms"], kernels[dom]["achieved"], kernels[dom]["traffic"]
        else:       # a candidate kernel of the first generation was selected (MXE_CAND_VARIANT < 4, or k / canonical outside variant 4)
            dom = "cand31_kernel"
            cand_ms = t_cand / max(1, n_cand)
            achieved = algo_bytes_per_launch / (cand_ms * 1e-3) / 1e9 if cand_ms > 0 else 0.0
            traffic = None
        # T1 (whole sketch per assembly): from the timed region = (step - steps 2-3) / assemblies, overlap included
        sketch_ms_per_asm = (sec / args.steps * 1e3 - t_filter_timed / args.steps) / n_asm
        pack_cand_ms_per_asm = (t_pack + t_cand) / args.steps / n_asm
        value = total_bases * args.steps / sec / 1e9
        line = {
            "metric": f"Gbases/s sketched+filtered at k={K} w={W}", "value": value, "unit": "Gbases/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": spec["name"], "k": K, "w": W, "bases_per_step": total_bases, "n_free": not args.with_n,
                       "l2": "inputs larger than L2 (>= 200 MB per assembly)",
                       "sketch_calls": "one call for all assemblies of the rank" if my_bases + (n_records + 8) * W < 0xF0000000 else
                                       ("one per assembly, enqueued on two streams (concurrent)" if overlap_mode == "1" else
                                        "one per assembly, enqueued back to back on the engine stream, sizes read once at the end" if overlap else
                                        "one per assembly, sequential"),
                       "sharding": f"contiguous record ranges over {world} rank(s)" +
                       (("; steps 2-3 by hash owner, 3 all-to-alls (NCCL): keys, marks, sightings" if dist_mode == "alltoall" else
                         "; steps 2-3 by hash range / own records, 1 all-gather + 3 all-reduces (NCCL)" if dist_mode == "allreduce" else
                         "; steps 2-3 by hash-bucket owner, exchanges as direct peer stores over NVLink (CUDA IPC) with device-side barriers, no collective library")
                        if world > 1 else ""),
                       "minimizers": stats.get("n_mx_total", stats["n_mx"]), "vertices": stats["vertices"], "edges": stats["edges"]},
            "e2e": {"value": total_bases * args.steps / sec_e2e / 1e9, "unit": "Gbases/s", "ms_per_step": sec_e2e / args.steps * 1e3,
                    "h2d_bytes_per_step": stats.get("h2d_total", my_bases), "d2h_bytes_per_step": stats.get("d2h_total", stats["d2h"])},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "launch_ms": cand_ms, "algorithmic_bytes_per_launch": algo_bytes_per_launch,
                         "kernels": kernels,
                         "pack_cand_frac": algo_bytes_per_launch / (pack_cand_ms_per_asm * 1e-3) / 1e9 / peak if pack_cand_ms_per_asm > 0 else None,
                         "sketch_frac": algo_bytes_per_launch / (sketch_ms_per_asm * 1e-3) / 1e9 / peak if sketch_ms_per_asm > 0 else None,
                         "sketch_ms_per_assembly": sketch_ms_per_asm, "pack_cand_ms_per_assembly": pack_cand_ms_per_asm,
                         "phase_ms_per_step": phases,
                         "phase_note": "phases and kernels timed in a separate pass with the assemblies sketched one after the other; "
                                       "sketch_frac is from the timed region"},
            "clocks": clocks,
        }
        alu_kernel = "scan_bs2_kernel" if "scan_bs2_kernel" in kernels else "cand31_kernel"
        line["roofline"]["alu"] = alu_roofline(my_bases / n_asm, kernels[alu_kernel]["launch_ms"] if alu_kernel in kernels else cand_ms, alu_kernel)
        if parity is not None:
            line["parity"] = parity
        if not args.no_cpu_baseline:
            cb, _ = cpu_reference_run(spec, args, 1, 1, with_t4=True)
            line["cpu_baseline"] = cb
        if world == 1 and not args.no_t3 and not args.no_cpu_baseline:
            try:
                del shards, combo, host
                torch.cuda.empty_cache()
                line["t3"] = run_t3(args, eng, spec, dev)
            except Exception as exc:           # reporting only: never lose the bench line over it
                line["t3"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["vs_single_gpu"]:
            print("bench.py: the merged multi-GPU result differs from the single-GPU result: " + ", ".join(parity["mismatch"]), file=sys.stderr, flush=True)
            parity_failed = True
    if p2p is not None:
        dist.barrier()
        p2p.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    if parity_failed:
        sys.exit(3)


if __name__ == "__main__":
    main()
