"""The C-ABI libraries load and export every symbol the headers declare (no compute, no GPU)."""
import ctypes, os, re
from conftest import ROOT
from checkers import Edge264Frame, _LIBS


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(edge264_\w+|e264b_\w+)\s*\(", txt)))


def test_frame_layout_matches_reference():
    # reference edge264.h:45-62 -> 96 bytes on LP64, offsets probed in SURVEY.md §8(b)
    assert ctypes.sizeof(Edge264Frame) == 96
    assert Edge264Frame.samples_mvc.offset == 24 and Edge264Frame.mb_errors.offset == 48
    assert Edge264Frame.bit_depth_Y.offset == 56 and Edge264Frame.width_Y.offset == 58
    assert Edge264Frame.stride_Y.offset == 66 and Edge264Frame.stride_mb.offset == 70
    assert Edge264Frame.FrameId.offset == 72 and Edge264Frame.frame_crop_offsets.offset == 80 and Edge264Frame.return_arg.offset == 88


def test_product_library_exports_both_headers():
    lib = ctypes.CDLL(_LIBS["gpu"])   # loading needs libcudart only, not a device
    names = declared("edge264.h") + declared("e264b_recon.h")
    assert len(names) >= 7 + 15
    for n in names:
        assert hasattr(lib, n), n


def test_oracle_library_exports_api():
    lib = ctypes.CDLL(_LIBS["port"])
    for n in declared("edge264.h"):
        assert hasattr(lib, n), n


def test_find_start_code_semantics():
    # reference edge264.c:87-119: pointer to the 00 00 01 (three-byte) or 00 00 00 01 (four-byte) prefix, else `end`
    from checkers import load
    for backend in ("port", "ref"):
        if not os.path.exists(_LIBS[backend]):
            continue
        lib = load(backend)
        data = bytes([9, 9, 0, 0, 1, 5, 0, 0, 0, 1, 7, 0, 0, 2, 0, 0]) + bytes(32)
        buf = ctypes.create_string_buffer(data, len(data)); base = ctypes.addressof(buf); end = base + 16
        assert lib.edge264_find_start_code(base, end, 0) - base == 2
        assert lib.edge264_find_start_code(base + 5, end, 0) - base == 7
        assert lib.edge264_find_start_code(base + 3, end, 1) - base == 6
        assert lib.edge264_find_start_code(base + 11, end, 0) == end
        assert lib.edge264_find_start_code(end, end, 0) == end
