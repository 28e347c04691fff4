"""The decoder's output logic with a backend whose pictures are "still on the device" for a while (oracle port backend,
E264_PORT_LATE=n: a picture counts as finished after n polls or a wait): edge264_get_frame answers ENOMSG instead of
waiting — the reference's answer for a frame its workers have not finished (edge264.c:373) — the application loop keeps
feeding NAL units, every frame still arrives, in the same order, and the decoder only WAITS for a picture after
edge264_decode_NAL has returned ENOBUFS or at the end of the stream."""
import ctypes, errno, hashlib, json, os, subprocess, sys
import pytest
from conftest import STREAMS, make_stream

CODE = r'''
import sys, json, ctypes, errno, hashlib
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[3])
from checkers import load, Edge264Frame
from edge264_b200 import _frame_bytes
lib = load("port")
raw = ctypes.CDLL(lib._name)
raw.e264_port_waits.restype = ctypes.c_int
data = open(sys.argv[1], "rb").read(); n_threads = int(sys.argv[4])
buf = ctypes.create_string_buffer(data, len(data) + 64); base = ctypes.addressof(buf); end = base + len(data)
dec = lib.edge264_alloc(n_threads, None, None, 0, None, None, None)
nal = base + 3 + (1 if data[2] == 0 else 0)
f = Edge264Frame(); frames = []; drained = False; bad_waits = 0
while True:
    sc = lib.edge264_find_start_code(nal, end, 0) if nal < end else end
    before = len(frames)
    res = lib.edge264_decode_NAL(dec, nal, sc, None, None)
    if nal >= end: drained = True
    w0 = raw.e264_port_waits()
    while True:
        r = lib.edge264_get_frame(dec, ctypes.byref(f), 0)
        if r != 0: break
        frames.append([f.FrameId, hashlib.md5(_frame_bytes(f)).hexdigest()])
    if res == 0 and not drained: bad_waits += raw.e264_port_waits() - w0      # between two NAL units get_frame must not wait
    if res == errno.ENOBUFS:
        if len(frames) == before: break
        continue
    nal = sc + 3 if sc + 3 < end else end
    if (res not in (0, errno.ENOTSUP, errno.EBADMSG)) or drained: break
d = ctypes.c_void_p(dec); lib.edge264_free(ctypes.byref(d))
print(json.dumps({"frames": frames, "bad_waits": bad_waits, "waits": raw.e264_port_waits()}))
'''


@pytest.mark.parametrize("name", ["b_explicit", "p_refs_wp", "dpb_mmco_cabac"])
@pytest.mark.parametrize("n_threads", [0, 2])
def test_late_pictures_arrive_completely_and_in_order(workdir, name, n_threads):
    nm, w, h, args = next(s for s in STREAMS if s[0] == name)
    path = make_stream(workdir, nm, w, h, args)
    here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
    out = {}
    for late in ("0", "1", "5", "1000000"):       # never late, late by one poll, by five, "never finishes unless waited for"
        r = subprocess.run([sys.executable, "-c", CODE, path, root, here, str(n_threads)], env=dict(os.environ, E264_PORT_LATE=late), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[late] = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(out["0"]["frames"]) > 0
    for late in ("1", "5", "1000000"):
        assert out[late]["frames"] == out["0"]["frames"]
        assert out[late]["bad_waits"] == 0
    assert out["1000000"]["waits"] > 0          # pictures that never finish by themselves were waited for at ENOBUFS / the end of the stream
