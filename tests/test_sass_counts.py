"""CPU: the per-base instruction counts DESIGN.md 3.3 and bench.py (roofline.alu, ALU_OPS_PER_BASE) quote for
cand31_kernel are read off the SASS of the library that ships -- this test re-derives them with cuobjdump so the quoted
figures cannot drift from the binary."""
import os
import re
import shutil
import subprocess

import pytest

import ntjoin_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALU = ("LOP3", "SHF", "PRMT", "ISETP", "IADD3", "VIADD", "LEA", "SEL", "VIMNMX", "VIADDMNMX", "POPC", "FLO", "BREV", "IABS")
FMA = ("IMAD", "FFMA", "FMUL")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOMS", "RED")


def main_loop(sass, outermost=False):
    """instructions of the innermost (or outermost) loop that holds sixteen LDS.64 (one group of 16 bases / steps)"""
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, text in ins:
        m = re.search(r"BRA\s+(?:`\(\S+\)|0x([0-9a-f]+))", text)
        if m and m.group(1) and int(m.group(1), 16) < addr:
            body = [t for a, t in ins if int(m.group(1), 16) <= a <= addr]
            if sum("LDS.64" in t for t in body) == 16 and (best is None or (len(body) > len(best) if outermost else len(body) < len(best))):
                best = body
    return best


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_cand31_instruction_mix_per_base():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    out = subprocess.run(["cuobjdump", "-sass", ntjoin_b200.library_path()], capture_output=True, text=True, check=True).stdout
    chunks = out.split("Function : ")
    fn = [c for c in chunks if c.startswith("_ZN3mxe13cand31_kernelILi0ELi1E")]
    assert len(fn) == 1, "cand31_kernel<0, 1> (sum combiner, FMA-offloaded test) not found in libmxe.so"
    body = main_loop(fn[0])
    assert body is not None, "main loop with 16 LDS.64 not found"
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]      # noqa: E731
    ops = [op(t) for t in body]
    alu = sum(o in ALU for o in ops) / 16.0
    fma = sum(o in FMA for o in ops) / 16.0
    lsu = sum(o in LSU for o in ops) / 16.0
    # 16 bases per trip: the roll is 5 ALU ops per base (2 rotations x 2 + ... ), test + accumulate on the FMA pipe
    assert sum(t.startswith("LDS.64") or " LDS.64" in t for t in body) == 16
    assert abs(alu - bench.ALU_OPS_PER_BASE["cand31_kernel"]) <= 0.25, (alu, fma, lsu)
    assert 4.0 <= fma <= 6.0 and 1.0 <= lsu <= 2.0, (alu, fma, lsu)
    assert len(ops) / 16.0 < 17.5            # whole trip, including loop overhead


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_scan_bs2_instruction_mix_per_base():
    """scan_bs2_kernel (k = 32): the loop over groups of 16 steps serves 16 x 32 = 512 positions per trip.  One LOP3 per
    state bit (62) + the base combos (10) + the 12-bit threshold adder (35) per step, i.e. ~113 LOP3 per 32 positions;
    a compiler that re-derives the combos inside the state updates costs 96 instead of 72 (scan_kernels.cuh)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    out = subprocess.run(["cuobjdump", "-sass", ntjoin_b200.library_path()], capture_output=True, text=True, check=True).stdout
    fn = [c for c in out.split("Function : ") if c.startswith("_ZN3mxe15scan_bs2_kernelILi1E")]
    assert len(fn) == 1, "scan_bs2_kernel<1, ...> (k = 32) not found in libmxe.so"
    assert "LDG.E.EFL2.256" in fn[0] or ".256" in fn[0], "the plane rows are read with 256-bit loads (sm_100)"
    body = main_loop(fn[0], outermost=True)
    assert body is not None, "group loop with 16 LDS.64 (the planes that leave the k-mer) not found"
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]      # noqa: E731
    ops = [op(t) for t in body]
    alu = sum(o in ALU for o in ops) / 512.0
    lop3 = sum(o == "LOP3" for o in ops) / 16.0
    assert abs(alu - bench.ALU_OPS_PER_BASE["scan_bs2_kernel"]) <= 0.2, alu
    assert 105 <= lop3 <= 120, lop3            # per step of 32 positions
    assert len(ops) / 512.0 < 4.4              # whole trip


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_steps23_kernels_keep_their_parameters_in_place():
    """Every kernel of csrc/p2p.cu indexes the table of peer pointers with a run-time rank.  Taking the table by value and
    passing it on by reference made the compiler copy all 128 bytes into local memory in every thread (16 STL.64 per
    thread; steps 2-3 took 2.8 instead of 1.7 ms per step); `const __grid_constant__` parameters are read where they are.
    The shipped SASS of these kernels must not touch local memory at all."""
    out = subprocess.run(["cuobjdump", "-sass", ntjoin_b200.library_path()], capture_output=True, text=True, check=True).stdout
    fns = [c for c in out.split("Function : ") if re.match(r"_ZN3mxe\d+p2p_", c)]
    assert len(fns) >= 20, "the kernels of csrc/p2p.cu were not found in libmxe.so"
    for fn in fns:
        name = fn.split()[0]
        assert not re.search(r"\b(STL|LDL)\b", fn), name + " spills or copies its parameters to local memory"


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_hash_tables_are_staged_by_bulk_copies():
    """cand_hash_pos_kernel and final_eval_pos_kernel stage their position-specific tables with cp.async.bulk against an
    mbarrier: UBLKCP + the transaction-count arrive + the phase wait must be in their SASS"""
    out = subprocess.run(["cuobjdump", "-sass", ntjoin_b200.library_path()], capture_output=True, text=True, check=True).stdout
    chunks = out.split("Function : ")
    for mangled in ("_ZN3mxe20cand_hash_pos_kernel", "_ZN3mxe21final_eval_pos_kernel"):
        fn = [c for c in chunks if c.startswith(mangled)]
        assert len(fn) == 1, mangled
        assert fn[0].count("UBLKCP.S.G") == 2, mangled
        assert "SYNCS.ARRIVE.TRANS64" in fn[0] and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in fn[0], mangled
