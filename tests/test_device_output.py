"""Zero-copy output: the application's alloc_cb may hand the decoder DEVICE memory; the frames then
come back as device pointers (same layout, same FrameIds) and never cross PCIe.  Checked against the host-mirror decode."""
import ctypes, errno, hashlib
import pytest
from conftest import STREAMS, make_stream
from checkers import load, Edge264Frame, decode_bytes

pytestmark = pytest.mark.gpu


def test_alloc_cb_with_device_memory_returns_device_frames(workdir):
    import torch
    name, w, h, args = next(s for s in STREAMS if s[0] == "b_explicit")
    data = open(make_stream(workdir, name, w, h, args), "rb").read()
    want = [(f[0], hashlib.md5(f[3]).hexdigest()) for f in decode_bytes(data, "gpu")[0]]
    lib = load("gpu")
    keep = []
    ALLOC = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint, ctypes.c_int, ctypes.c_void_p)
    FREE = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)

    def alloc(samples, samples_size, mbs, mbs_size, err, arg):
        t = torch.empty(samples_size + 256, dtype=torch.uint8, device="cuda")
        keep.append(t)
        samples[0] = t.data_ptr(); mbs[0] = None
    alloc_cb, free_cb = ALLOC(alloc), FREE(lambda a, b, c: None)
    lib.edge264_alloc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ALLOC, FREE, ctypes.c_void_p]
    dec = lib.edge264_alloc(0, None, None, 0, alloc_cb, free_cb, None)
    assert dec
    buf = ctypes.create_string_buffer(data, len(data) + 64); base = ctypes.addressof(buf); end = base + len(data)
    nal = base + 4; f = Edge264Frame(); got = []; drained = False
    while True:
        sc = lib.edge264_find_start_code(nal, end, 0) if nal < end else end
        before = len(got)
        res = lib.edge264_decode_NAL(dec, nal, sc, None, None)
        if nal >= end: drained = True
        while lib.edge264_get_frame(dec, ctypes.byref(f), 0) == 0:
            out = bytearray()
            for pl in range(3):
                wd, ht, st = (f.width_C, f.height_C, f.stride_C) if pl else (f.width_Y, f.height_Y, f.stride_Y)
                ptr = ctypes.cast(f.samples[pl], ctypes.c_void_p).value
                owner = next(t for t in keep if t.data_ptr() <= ptr < t.data_ptr() + t.numel())       # a device pointer inside one of OUR tensors
                off = ptr - owner.data_ptr()
                plane = owner[off:off + st * ht].view(ht, st)[:, :wd].contiguous().cpu().numpy().tobytes()
                out += plane
            got.append((f.FrameId, hashlib.md5(bytes(out)).hexdigest()))
        if res == errno.ENOBUFS:
            if len(got) == before: break
            continue
        nal = sc + 3 if sc + 3 < end else end
        if (res not in (0, errno.ENOTSUP, errno.EBADMSG)) or drained: break
    d = ctypes.c_void_p(dec); lib.edge264_free(ctypes.byref(d))
    lib.edge264_alloc.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    assert got == want and len(keep) > 0
