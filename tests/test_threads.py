"""edge264_alloc(n_threads != 0): slice data is parsed ahead on worker threads (reference: slice tasks + worker_loop,
/root/reference/src/edge264.c:223-257, edge264_headers.c:450-603).  Output — FrameIds, order, every sample — must be what
the synchronous decoder produces; the CPU legs run the product's host sources over the checker backend, the GPU leg the
product library."""
import hashlib, os
import pytest
from conftest import STREAMS, DPB_STREAMS, make_stream
from checkers import decode_bytes

PICK = ["i_cavlc_slices", "p_refs_wp", "b_implicit_temporal", "b_cavlc_all", "dpb_mmco_cabac", "bref_spatial",
        "dpb_ipb_temporal", "mixed_slices_cabac", "direct4x4_spatial", "nonref_p"]
CASES = [s for s in STREAMS + DPB_STREAMS if s[0] in PICK]


def digest(frames):
    return [(f[0], hashlib.md5(f[3]).hexdigest()) for f in frames]


@pytest.mark.parametrize("name,w,h,args", CASES, ids=[c[0] for c in CASES])
def test_parse_ahead_equals_synchronous(workdir, name, w, h, args):
    data = open(make_stream(workdir, name, w, h, args), "rb").read()
    want, codes = decode_bytes(data, "port", 0)
    for n_threads in (1, 3, 4):
        got, _ = decode_bytes(data, "port", n_threads)
        assert digest(got) == digest(want), "n_threads=%d" % n_threads


def test_flush_and_free_with_pictures_in_flight(workdir):
    """edge264_flush / edge264_free while workers still hold pictures: nothing may hang or leak into the next sequence."""
    import ctypes
    from checkers import load, Edge264Frame
    lib = load("port")
    data = open(make_stream(workdir, "thr_flush", 9, 7, "-n 12 -s 77 --gop IPB --deblock 0 --refs 3"), "rb").read()
    buf = ctypes.create_string_buffer(data, len(data) + 64); base = ctypes.addressof(buf); end = base + len(data)
    for cut in (3, 7, 11):
        dec = lib.edge264_alloc(3, None, None, 0, None, None, None)
        nal, k = base + 4, 0
        while nal < end and k < cut:
            sc = lib.edge264_find_start_code(nal, end, 0)
            lib.edge264_decode_NAL(dec, nal, sc, None, None)
            nal = sc + 3; k += 1
        if cut == 7:
            lib.edge264_flush(dec)
        d = ctypes.c_void_p(dec); lib.edge264_free(ctypes.byref(d))
    want, _ = decode_bytes(data, "port", 0)
    got, _ = decode_bytes(data, "port", 3)
    assert digest(got) == digest(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h,args", [c for c in CASES if c[0] in ("b_implicit_temporal", "bref_spatial", "mixed_slices_cabac")] +
                         [("thr_1080p", 120, 68, "-n 13 -s 32 --gop IPB --deblock 0 --t8x8 50 --density 52 --wp 2")],
                         ids=["b_implicit_temporal", "bref_spatial", "mixed_slices_cabac", "thr_1080p"])
def test_gpu_parse_ahead(workdir, name, w, h, args):
    data = open(make_stream(workdir, name, w, h, args), "rb").read()
    want, _ = decode_bytes(data, "gpu", 0)
    for n_threads in (2, 4):
        got, _ = decode_bytes(data, "gpu", n_threads)
        assert digest(got) == digest(want), "n_threads=%d" % n_threads
