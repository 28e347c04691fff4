"""Regenerates tests/golden/streams.json from the REFERENCE decoder (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container: python tests/make_golden.py"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT, STREAMS, DPB_STREAMS, make_stream, md5_frames
from checkers import decode_bytes

out = {}
tmp = tempfile.mkdtemp()
for name, w, h, args in STREAMS + DPB_STREAMS:
    data = open(make_stream(tmp, name, w, h, args), "rb").read()
    frames, _ = decode_bytes(data, "ref")
    out[name] = {"args": f"-W {w} -H {h} {args}", "bytes": len(data), "md5": md5_frames(frames)}
refs = {}
for f in ["finish-frame.264", "nal-ref-idc-0.264", "non-ref-dec-poc.264", "poc-out-of-order.264", "pos-frame-num-idr.264", "supp-nals.264", "zero-cropping.264"]:
    frames, _ = decode_bytes(open("/root/reference/tests/" + f, "rb").read(), "ref")
    refs[f] = md5_frames(frames)
out["_reference_streams"] = refs
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "streams.json"), "w"), indent=1)
print("wrote", len(out) - 1, "streams")
