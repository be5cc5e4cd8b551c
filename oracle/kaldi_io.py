"""Minimal Kaldi binary matrix / vector file I/O for fixtures (formats: src/matrix/kaldi-matrix.cc:1201-1227,
kaldi-vector.cc:1210-1230; file header "\\0B", src/util/kaldi-io.cc).  TEST INFRASTRUCTURE ONLY."""
import struct

import numpy as np


def write_mat(path, arr):
    a = np.ascontiguousarray(arr, np.float32)
    with open(path, "wb") as f:
        f.write(b"\0BFM \x04" + struct.pack("<i", a.shape[0]) + b"\x04" + struct.pack("<i", a.shape[1]))
        f.write(a.tobytes())


def write_vec(path, arr):
    a = np.ascontiguousarray(arr, np.float32)
    with open(path, "wb") as f:
        f.write(b"\0BFV \x04" + struct.pack("<i", a.shape[0]))
        f.write(a.tobytes())


def _read(f):
    tok = f.read(3)
    if tok == b"FM ":
        assert f.read(1) == b"\x04"
        r = struct.unpack("<i", f.read(4))[0]
        assert f.read(1) == b"\x04"
        c = struct.unpack("<i", f.read(4))[0]
        return np.frombuffer(f.read(4 * r * c), np.float32).reshape(r, c).copy()
    if tok == b"FV ":
        assert f.read(1) == b"\x04"
        n = struct.unpack("<i", f.read(4))[0]
        return np.frombuffer(f.read(4 * n), np.float32).copy()
    raise ValueError("unsupported token %r" % tok)


def read(path):
    with open(path, "rb") as f:
        assert f.read(2) == b"\0B", "not a Kaldi binary file: %s" % path
        return _read(f)
