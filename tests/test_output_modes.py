"""edge264_get_frame and the device (behavioural notes): a picture that is still being reconstructed
answers ENOMSG — what the reference answers for a frame its workers have not finished (edge264.c:373) — and the call
waits only after ENOBUFS / at the end of the stream.  The frames an application collects are the same, in the same order,
as with a get_frame that always waits (E264_SYNC_OUTPUT=1), and nothing is lost at the end of the stream."""
import hashlib, json, os, subprocess, sys
import pytest
from conftest import STREAMS, make_stream

pytestmark = pytest.mark.gpu

CODE = r'''
import sys, json, hashlib
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, sys.argv[3])
from checkers import decode_bytes
frames, codes = decode_bytes(open(sys.argv[1], "rb").read(), "gpu", int(sys.argv[4]))
print(json.dumps([[f[0], hashlib.md5(f[3]).hexdigest()] for f in frames]))
'''


@pytest.mark.parametrize("name", ["b_explicit", "p_refs_wp"])
@pytest.mark.parametrize("n_threads", [0, 2])
def test_polling_and_waiting_get_frame_deliver_the_same_frames(workdir, name, n_threads):
    cand = [s for s in STREAMS if s[0] == name] or [STREAMS[0]]
    nm, w, h, args = cand[0]
    path = make_stream(workdir, nm, w, h, args)
    here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
    out = {}
    for sync in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", CODE, path, root, here, str(n_threads)], env=dict(os.environ, E264_SYNC_OUTPUT=sync), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[sync] = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["0"] == out["1"] and len(out["0"]) > 0
    port = subprocess.run([sys.executable, "-c", CODE.replace('"gpu"', '"port"'), path, root, here, "0"], capture_output=True, text=True, timeout=300)
    assert json.loads(port.stdout.strip().splitlines()[-1]) == out["0"]
