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


# ----------------------------------------------------------------------------- binary <Nnet> model files
class _Stream:
    def __init__(self, data):
        self.d, self.p = data, 0

    def token(self):
        e = self.d.index(b" ", self.p)
        t = self.d[self.p:e].decode()
        self.p = e + 1
        return t

    def peek(self):
        return self.d[self.p:self.p + 1]

    def int32(self):
        assert self.d[self.p] == 4
        v = struct.unpack_from("<i", self.d, self.p + 1)[0]
        self.p += 5
        return v

    def float32(self):
        assert self.d[self.p] == 4
        v = struct.unpack_from("<f", self.d, self.p + 1)[0]
        self.p += 5
        return v

    def float64(self):
        assert self.d[self.p] == 8
        v = struct.unpack_from("<d", self.d, self.p + 1)[0]
        self.p += 9
        return v

    def intvec(self):
        assert self.d[self.p] == 4
        n = struct.unpack_from("<i", self.d, self.p + 1)[0]
        self.p += 5
        v = list(struct.unpack_from("<%di" % n, self.d, self.p))
        self.p += 4 * n
        return v

    def mat(self):
        tok = self.token()
        assert tok in ("FM", "DM"), tok
        r, c = self.int32(), self.int32()
        dt, sz = (np.float32, 4) if tok == "FM" else (np.float64, 8)
        a = np.frombuffer(self.d, dt, r * c, self.p).reshape(r, c).copy()
        self.p += sz * r * c
        return a

    def vec(self):
        tok = self.token()
        assert tok in ("FV", "DV"), tok
        n = self.int32()
        dt, sz = (np.float32, 4) if tok == "FV" else (np.float64, 8)
        a = np.frombuffer(self.d, dt, n, self.p).copy()
        self.p += sz * n
        return a


def _lstm_dir(s, projected):
    d = {"w_x": s.mat(), "w_r": s.mat(), "bias": s.vec(), "pi": s.vec(), "pf": s.vec(), "po": s.vec()}
    if projected:
        d["w_rm"] = s.mat()
    return d


def read_nnet(path):
    """Parses a binary model file written by Nnet::Write (src/aslp-nnet/nnet-nnet.cc:644-653, nnet-component.cc:328-342 and
    the WriteData of each component) into a list of dicts.  Only the component types of the hot path."""
    data = open(path, "rb").read()
    assert data[:2] == b"\0B"
    s = _Stream(data)
    s.p = 2
    comps = []
    tok = s.token()
    assert tok == "<Nnet>"
    while True:
        tok = s.token()
        if tok == "</Nnet>":
            break
        c = {"type": tok.strip("<>"), "out_dim": s.int32(), "in_dim": s.int32()}
        if s.peek() == b"<":
            assert s.token() == "<Name>"
            c["name"] = s.token()
        c["id"], c["input"], c["offset"] = s.int32(), s.intvec(), s.intvec()
        t = c["type"]
        if t == "AffineTransform":
            assert s.token() == "<LearnRateCoef>"; c["lr_coef"] = s.float32()
            assert s.token() == "<BiasLearnRateCoef>"; c["bias_lr_coef"] = s.float32()
            assert s.token() == "<MaxNorm>"; c["max_norm"] = s.float32()
            c["W"], c["b"] = s.mat(), s.vec()
        elif t == "LinearTransform":
            assert s.token() == "<LearnRateCoef>"; c["lr_coef"] = s.float32()
            c["W"] = s.mat()
        elif t in ("Lstm", "BLstm"):
            assert s.token() == "<ClipGradient>"; c["clip"] = s.float32()
            c["dirs"] = [_lstm_dir(s, False) for _ in range(2 if t == "BLstm" else 1)]
        elif t in ("LstmProjectedStreams", "BLstmProjectedStreams", "BLstmProjectedStreamsLC"):
            assert s.token() == "<CellDim>"; c["cell"] = s.int32()
            assert s.token() == "<ClipGradient>"; c["clip"] = s.float32()
            c["dirs"] = [_lstm_dir(s, True) for _ in range(1 if t == "LstmProjectedStreams" else 2)]
        elif t == "GruStreams":
            assert s.token() == "<ClipGradient>"; c["clip"] = s.float32()
            c["w_zrm_x"], c["w_zr_h"], c["w_m_g"], c["bias"] = s.mat(), s.mat(), s.mat(), s.vec()
        elif t == "CompactFsmn":
            assert s.token() == "<PastContext>"; c["past"] = s.int32()
            assert s.token() == "<FutureContext>"; c["future"] = s.int32()
            assert s.token() == "<LearnRateCoef>"; c["lr_coef"] = s.float32()
            c["coef"] = s.mat()
        elif t == "BatchNormalization":
            assert s.token() == "<NumAccFrames>"; c["num_acc_frames"] = s.float64()
            c["acc_means"], c["acc_vars"], c["shift"], c["scale"] = s.vec(), s.vec(), s.vec(), s.vec()
        elif t == "Splice":
            c["offsets"] = s.intvec()
        elif t == "RowConvolution":
            assert s.token() == "<FutureContext>"; c["future"] = s.int32()
            c["w"] = s.mat()
        elif t == "ConvolutionalComponent":      # nnet-convolutional-component.h:199-221
            assert s.token() == "<PatchDim>"; c["patch_dim"] = s.int32()
            assert s.token() == "<PatchStep>"; c["patch_step"] = s.int32()
            assert s.token() == "<PatchStride>"; c["patch_stride"] = s.int32()
            assert s.token() == "<LearnRateCoef>"; c["lr_coef"] = s.float32()
            assert s.token() == "<BiasLearnRateCoef>"; c["bias_lr_coef"] = s.float32()
            assert s.token() == "<MaxNorm>"; c["max_norm"] = s.float32()
            assert s.token() == "<Filters>"; c["filters"] = s.mat()
            assert s.token() == "<Bias>"; c["bias"] = s.vec()
        elif t == "MaxPoolingComponent":         # nnet-max-pooling-component.h:91-98
            assert s.token() == "<PoolSize>"; c["pool_size"] = s.int32()
            assert s.token() == "<PoolStep>"; c["pool_step"] = s.int32()
            assert s.token() == "<PoolStride>"; c["pool_stride"] = s.int32()
        elif t in ("Softmax", "Sigmoid", "Tanh", "ReLU", "InputLayer", "OutputLayer"):
            pass
        else:
            raise ValueError("unsupported component in fixture: " + t)
        comps.append(c)
    return comps
