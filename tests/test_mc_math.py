"""edge264_b200/csrc/mc_math.cuh (the per-thread 4x4 interpolation of the inter kernel) compiles for the host too:
tests/native/mc_math_check.cpp runs it over every fractional position on random and extreme windows against the oracle's
per-sample restatement (oracle/port_recon.c, pinned to the reference's decode_inter_luma by `ref_kat fuzz`)."""
import os, subprocess
from conftest import ROOT


def test_register_interpolation_matches_the_oracle(tmp_path):
    obj, exe = str(tmp_path / "port_recon.o"), str(tmp_path / "mc_math_check")
    subprocess.run(["gcc", "-O2", "-w", "-c", os.path.join(ROOT, "oracle", "port_recon.c"), "-o", obj], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "native", "mc_math_check.cpp"), obj, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-500:]
