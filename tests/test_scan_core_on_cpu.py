"""CPU: the per-thread core of scan_bs2_kernel (bs2_tile: bit-plane input, 62-bit bit-sliced roll, phase bookkeeping,
ring of leaving bases, 12-bit threshold test, emission queue) and the byte-validity screen / packing / 32x32 bit
transposition of pack2_kernel are `__host__ __device__` in ntjoin_b200/csrc/scan_kernels.cuh.  tools/scan_emul.cu runs
exactly that source on the host against plain integer arithmetic (every candidate a superset member, every k-mer below
the threshold flagged, no position emitted twice, all 256 byte values in all 32 slots).  This test compiles and runs it:
the shipped kernel source is exercised in the CPU tier, not only its restatements."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_scan_kernel_core_and_pack_screen_on_the_host(tmp_path):
    exe = str(tmp_path / "scan_emul")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-w",
                    "-I", os.path.join(ROOT, "ntjoin_b200", "csrc"), "-o", exe, os.path.join(ROOT, "tools", "scan_emul.cu")],
                   check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert r.returncode == 0, r.stdout
    assert len(lines) >= 9 and all(l.endswith("ok") for l in lines), r.stdout
    assert any("k=40" in l for l in lines) and any("k=24" in l for l in lines)
