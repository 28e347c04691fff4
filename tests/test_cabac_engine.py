"""The CABAC bin decoder of the host parser (edge264_b200/csrc/cabac.h) against a literal restatement of 9.3.3.2 on
bins written by the encoder of 9.3.4.2 (tests/native/cabac_check.c): bin values and context states after every bin."""
import os, subprocess
from conftest import ROOT


def test_bin_decoder_matches_the_standard_step_by_step(tmp_path):
    exe = str(tmp_path / "cabac_check")
    subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-std=gnu11", "-w", "-DE264_ENCODER", os.path.join(ROOT, "tests", "native", "cabac_check.c"), "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout[-800:]
