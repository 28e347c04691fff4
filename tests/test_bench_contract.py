"""bench.py's reference arm needs no GPU: run it on a tiny batch and check the one-JSON-line contract
(metric/unit/value, impl, cpu_baseline, e2e) the driver parses.  The GPU arm's line is checked on the B200."""
import json, os, subprocess, sys
import pytest
from conftest import ROOT


def test_reference_arm_prints_one_json_line(tmp_path):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libe264bench_ref.so")):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--streams", "2", "--frames", "6", "--workdir", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p_high_cabac_ipb_decode_fps" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
