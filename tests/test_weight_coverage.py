"""SURVEY config 5 / reference edge264_inter.c:1152-1175: the weighting corner cases must actually occur in the streams
that claim to cover them.  The oracle counts the 4x4 blocks predicted with implicit weights -64 / 128 and with the explicit
weight 128 (the inferred default at log2 denominator 7, uni-prediction only: 128 in bi-prediction is not a legal stream);
the same streams run bit-exact against the reference on the CPU and on the GPU (conftest STREAMS)."""
import os, re, subprocess
from conftest import ROOT, STREAMS, make_stream


def coverage(workdir, name):
    s = next(x for x in STREAMS if x[0] == name)
    path = make_stream(workdir, *s)
    r = subprocess.run([os.path.join(ROOT, "oracle", "oracle_decode"), path, "-q"], capture_output=True, text=True, env=dict(os.environ, E264_COVERAGE="1"), timeout=120)
    m = re.search(r"coverage implicit_w1_-64 (\d+) implicit_w1_128 (\d+) explicit_w128_uni (\d+) explicit_w128_bi (\d+)", r.stderr)
    assert m, r.stderr[-300:]
    return [int(x) for x in m.groups()]


def test_implicit_weight_extremes_are_reached(workdir):
    m64, p128, _, _ = coverage(workdir, "wp_implicit_extremes")
    assert m64 > 0 and p128 > 0


def test_explicit_weight_128_is_reached(workdir):
    _, _, uni, _ = coverage(workdir, "b_explicit")
    assert uni > 0
