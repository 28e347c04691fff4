"""Crafted header fields that once broke the host side (round-1 advisor findings): Exp-Golomb codes >= 2^31 landing in
ints (first_mb, pps_id, slice_type ...), 33 reference-list modifications (stack overflow in build_ref_lists), and a
format change while the application still borrows a frame (use after free).  The reference bounds every such read
(get_ue16/get_ue32 with a maximum, /root/reference/src/edge264_bitstream.c:150-203) and refuses the new SPS with
ENOBUFS while frames are out (edge264_headers.c:2005-2007).  Each case runs in a child process over the CPU checker
build of the product's host sources (oracle/liboracle_dec.so): a crash is a failure, any errno is fine."""
import errno, os, subprocess, sys
import pytest
from conftest import ROOT, make_stream


class Bits:
    def __init__(self): self.b = []
    def u(self, n, v): self.b += [(v >> i) & 1 for i in range(n - 1, -1, -1)]; return self
    def ue(self, v):
        x = v + 1; n = x.bit_length() - 1
        return self.u(n, 0).u(1, 1).u(n, x & ((1 << n) - 1)) if n else self.u(1, 1)
    def nal(self, ref_idc, typ):
        bits = self.b + [1]
        bits += [0] * (-len(bits) % 8)
        body = bytes(int("".join(map(str, bits[i:i + 8])), 2) for i in range(0, len(bits), 8))
        out, zeros = bytearray(), 0
        for c in body:
            if zeros >= 2 and c <= 3: out.append(3); zeros = 0
            out.append(c); zeros = zeros + 1 if c == 0 else 0
        return b"\0\0\0\1" + bytes([(ref_idc << 5) | typ]) + bytes(out)


def p_slice_prefix(first_mb=0, slice_type=0, pps_id=0):
    # gen264 streams without --dpb: log2_max_frame_num 8, poc type 0 with 10 lsb bits (tools/gen264.c:570)
    return Bits().ue(first_mb).ue(slice_type).ue(pps_id).u(8, 1).u(10, 2)


CHILD = r"""
import sys, ctypes
sys.path.insert(0, %r); sys.path.insert(0, %r)
from checkers import decode_bytes
frames, codes = decode_bytes(open(sys.argv[1], 'rb').read(), 'port')
print('codes', codes)
"""


def run_child(path):
    code = CHILD % (ROOT, os.path.join(ROOT, "tests"))
    return subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, timeout=120)


@pytest.fixture(scope="module")
def base(workdir):
    src = open(make_stream(workdir, "hard_base", 4, 3, "-n 3 -s 5 --gop IP --refs 2 --deblock 0"), "rb").read()
    nals = src.split(b"\0\0\0\1")[1:]
    first_p = next(i for i, n in enumerate(nals) if (n[0] & 31) == 1)
    return b"".join(b"\0\0\0\1" + n for n in nals[:first_p])   # SPS, PPS and the IDR picture


CASES = {
    "first_mb_2^31": p_slice_prefix(first_mb=0x80000000),
    "first_mb_max": p_slice_prefix(first_mb=0xfffffffe),
    "slice_type_2^31": p_slice_prefix(slice_type=0x80000005),
    "pps_id_2^31": p_slice_prefix(pps_id=0x80000000),
    "pps_id_big": p_slice_prefix(pps_id=200),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oversized_exp_golomb_fields_are_rejected(workdir, base, name):
    path = os.path.join(workdir, "hard_" + name.replace("^", "") + ".264")
    open(path, "wb").write(base + CASES[name].u(32, 0xdeadbeef).nal(2, 1))
    r = run_child(path)
    assert r.returncode == 0, r.stderr[-500:]
    codes = eval(r.stdout.split("codes", 1)[1])
    assert codes[-2] in (errno.EBADMSG, errno.ENOTSUP), codes   # the crafted slice; the last code is the flush


def test_33_list_modifications_do_not_overflow(workdir, base):
    b = p_slice_prefix().u(1, 1).ue(1).u(1, 1)      # num_ref_idx_active_override: 2 entries; ref_pic_list_modification_flag_l0
    for _ in range(33): b.ue(0).ue(0)
    b.ue(3).u(32, 0)
    path = os.path.join(workdir, "hard_rplm33.264")
    open(path, "wb").write(base + b.nal(2, 1))
    r = run_child(path)
    assert r.returncode == 0, r.stderr[-500:]


def test_sps_bounds(workdir, base):
    # width 2^31 macroblocks, 2^31 reference frames, log2_max_frame_num 2^31
    for k, sps in enumerate([Bits().u(8, 100).u(8, 0).u(8, 40).ue(0).ue(1).ue(0).ue(0).u(1, 0).u(1, 0).ue(0x7ffffffe),
                             Bits().u(8, 66).u(8, 0).u(8, 40).ue(0).ue(0).ue(2).ue(0x80000000).u(1, 0).ue(3).ue(3),
                             Bits().u(8, 66).u(8, 0).u(8, 40).ue(0).ue(0).ue(2).ue(1).u(1, 0).ue(0x80000000).ue(0x80000001)]):
        path = os.path.join(workdir, "hard_sps%d.264" % k)
        open(path, "wb").write(sps.u(16, 0xffff).nal(3, 7) + base)
        r = run_child(path)
        assert r.returncode == 0, r.stderr[-500:]


def test_borrowed_frame_survives_a_format_change(workdir):
    """A frame taken with borrow=1 must stay readable when an SPS with another size arrives: decode_NAL answers
    ENOBUFS until it is returned (reference bump_all_frames / to_get_frames test)."""
    a = open(make_stream(workdir, "hard_fmt_a", 4, 3, "-n 2 -s 6 --gop I --deblock 0"), "rb").read()
    b = open(make_stream(workdir, "hard_fmt_b", 6, 4, "-n 2 -s 7 --gop I --deblock 0"), "rb").read()
    code = r"""
import sys, ctypes, errno
sys.path.insert(0, %r); sys.path.insert(0, %r)
from checkers import load, Edge264Frame
lib = load('port')
data = open(sys.argv[1], 'rb').read() + open(sys.argv[2], 'rb').read()
buf = ctypes.create_string_buffer(data, len(data) + 64); base = ctypes.addressof(buf); end = base + len(data)
dec = lib.edge264_alloc(0, None, None, 0, None, None, None)
nal = base + 4; f = Edge264Frame(); held = None; saw_enobufs = False; snapshot = None
while nal < end:
    sc = lib.edge264_find_start_code(nal, end, 0)
    res = lib.edge264_decode_NAL(dec, nal, sc, None, None)
    if held is None and lib.edge264_get_frame(dec, ctypes.byref(f), 1) == 0:
        held = f.return_arg; snapshot = ctypes.string_at(f.samples[0], 64)
    if res == errno.ENOBUFS:
        if held is not None and held != 0:
            saw_enobufs = True
            assert ctypes.string_at(f.samples[0], 64) == snapshot      # still mapped, still the same samples
            lib.edge264_return_frame(dec, held); held = 0
            continue
        g = Edge264Frame()
        if lib.edge264_get_frame(dec, ctypes.byref(g), 0) != 0: break
        continue
    nal = sc + 3 if sc + 3 < end else end
print('enobufs', saw_enobufs)
""" % (ROOT, os.path.join(ROOT, "tests"))
    pa, pb = os.path.join(workdir, "hard_fmt_a.264"), os.path.join(workdir, "hard_fmt_b.264")
    r = subprocess.run([sys.executable, "-c", code, pa, pb], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-800:]
    assert "enobufs True" in r.stdout, r.stdout
