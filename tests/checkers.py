"""Test-side loaders: the product library next to the CPU checkers (oracle/liboracle_dec.so = the product's host
sources over the plain-C restatement; oracle/_ref/libedge264_ref.so = the unmodified reference compiled here).
Only tests, __graft_entry__.smoke() and the tools under tools/ import this; the shipped package knows nothing of oracle/."""
import os, hashlib
import edge264_b200 as _pkg
from edge264_b200 import Edge264Frame, _frame_bytes, ROOT  # noqa: F401  (re-exported for the tests)

_LIBS = {
    "gpu": _pkg.LIB_PATH,
    "port": os.path.join(ROOT, "oracle", "liboracle_dec.so"),
    "ref": os.path.join(ROOT, "oracle", "_ref", "libedge264_ref.so"),
}


def load(backend="gpu"):
    return _pkg.bind(_LIBS[backend])


def decode_bytes(data, backend="gpu", n_threads=0):
    return _pkg.decode_bytes(data, None if backend == "gpu" else load(backend), n_threads)


def decode_file_hashes(path, backend="gpu"):
    frames, _ = decode_bytes(open(path, "rb").read(), backend)
    return [hashlib.md5(fr[3]).hexdigest() for fr in frames]
