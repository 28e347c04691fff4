"""BASELINE.json configs[0]: a 1080p all-I High-profile CAVLC clip written by the REFERENCE's own generator
(/root/reference/tests/gen_avc.py; tests/make_gen_avc_fixture.py made it, tests/golden/gen_avc_1080p_intra_cavlc.json holds
the reference decoder's digests).  The only pixel-bearing stream of the suite whose bits were not written by the
repository's shared syntax code, so a misunderstanding shared by tools/gen264 and the parser cannot hide in it."""
import hashlib, json, os
import pytest
from conftest import ROOT
from checkers import decode_bytes

FIX = os.path.join(ROOT, "tests", "golden", "gen_avc_1080p_intra_cavlc")
GOLD = json.load(open(FIX + ".json"))


def digests(backend, n_threads=0):
    frames, codes = decode_bytes(open(FIX + ".264", "rb").read(), backend, n_threads)
    return [hashlib.md5(f[3]).hexdigest() for f in frames], [f[0] for f in frames]


def test_fixture_is_what_the_reference_decoded():
    assert os.path.getsize(FIX + ".264") == GOLD["bytes"]


def test_oracle_port_matches_reference_digests():
    md5, ids = digests("port")
    assert md5 == GOLD["md5"] and ids == GOLD["frame_ids"]


@pytest.mark.gpu
def test_gpu_matches_reference_digests():
    md5, ids = digests("gpu")
    assert md5 == GOLD["md5"] and ids == GOLD["frame_ids"]
    md5t, _ = digests("gpu", 3)
    assert md5t == GOLD["md5"]
