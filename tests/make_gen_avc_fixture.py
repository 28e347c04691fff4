"""Regenerates tests/golden/gen_avc_1080p_intra_cavlc.264 (+ .json): BASELINE.json configs[0], "1080p all-I High Profile
clip (synthetic via tests/gen_avc.py)".  Every bit of the fixture is written by the REFERENCE's own generator
/root/reference/tests/gen_avc.py (which turns a YAML description — the format of the reference decoder's log — into
Annex-B); the description is the reference decoder's log of a clip from tools/gen264, so the content is ours, the
bitstream writer is not: the one stream of the suite whose syntax the repository's shared parser/writer code never touched.
Run in the build container (needs /root/reference): python tests/make_gen_avc_fixture.py"""
import hashlib, json, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT
from checkers import decode_bytes

tmp = tempfile.mkdtemp()
src, yml, out = os.path.join(tmp, "a.264"), os.path.join(tmp, "a.yaml"), os.path.join(ROOT, "tests", "golden", "gen_avc_1080p_intra_cavlc.264")
args = "-W 120 -H 68 -n 2 -s 31 --gop I --deblock 0 --cavlc --t8x8 50 --density 14 --pcm 0"   # gen_avc.py cannot write I_PCM samples
subprocess.run([os.path.join(ROOT, "tools", "gen264"), "-o", src] + args.split(), check=True, stderr=subprocess.DEVNULL)
subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_log"), src, yml], check=True)
subprocess.run([sys.executable, "/root/reference/tests/gen_avc.py", yml, out], check=True)
frames, _ = decode_bytes(open(out, "rb").read(), "ref")
json.dump({"made_by": "/root/reference/tests/gen_avc.py from the reference decoder's log of: gen264 " + args, "bytes": os.path.getsize(out),
           "md5": [hashlib.md5(f[3]).hexdigest() for f in frames], "frame_ids": [f[0] for f in frames]},
          open(out.replace(".264", ".json"), "w"), indent=1)
print("wrote", out, os.path.getsize(out), "bytes,", len(frames), "frames")
