"""kaldi-aslp_b200: B200-native backing for the aslp-nnet training step.

This Python package is plumbing only (ctypes loaders for the two in-tree shared
libraries and a few helpers for tests / bench.py).  The product is native:

  libaslp_b200.so   hand-written sm_100a CUDA kernels behind include/aslp_b200.h, include/ctc.h
  libaslp_nnet.so   C++ Component / Nnet / IWorker mirror of src/aslp-nnet + src/aslp-parallel

There is no CPU fallback: loading fails loudly if the libraries are missing, and every
compute call fails without a CUDA device.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
LIB_CUDA_PATH = os.environ.get("ASLP_B200_CUDA_LIB", os.path.join(HERE, "libaslp_b200.so"))     # override: A/B builds of one kernel (tools/)
LIB_HOST_PATH = os.path.join(HERE, "libaslp_nnet.so")

_CTYPE = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
    "unsigned long long": ctypes.c_ulonglong, "aslp_stream_t": ctypes.c_void_p, "aslp_comm_t": ctypes.c_void_p,
    "aslp_nnet_t": ctypes.c_void_p, "aslp_worker_t": ctypes.c_void_p, "aslp_server_t": ctypes.c_void_p, "long long": ctypes.c_longlong,
    "ctcStatus_t": ctypes.c_int,
}


class CtcComputeInfo(ctypes.Structure):
    """struct ctcComputeInfo of include/ctc.h (loc + union{num_threads, stream})."""
    _fields_ = [("loc", ctypes.c_int), ("stream", ctypes.c_void_p)]


def parse_header(path):
    """Very small C declaration parser: returns {name: (restype, [argtypes])} for the
    function declarations of one of our headers (plain C, one declaration per ';')."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", " ", txt)
    txt = re.sub(r"#[^\n]*", " ", txt)
    # drop struct / enum / typedef bodies
    txt = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", txt, flags=re.S)
    txt = re.sub(r"typedef\s+enum\s*\{.*?\}\s*\w+\s*;", " ", txt, flags=re.S)
    txt = re.sub(r"struct\s+\w+\s*\{.*?\}\s*;", " ", txt, flags=re.S)
    txt = re.sub(r"enum\s*\{.*?\}\s*;", " ", txt, flags=re.S)
    txt = re.sub(r"typedef[^;]*;", " ", txt)
    txt = txt.replace('extern "C" {', " ").replace("}", " ")
    out = {}
    for decl in txt.split(";"):
        m = re.match(r"\s*([\w\s\*]+?)\s*\b(\w+)\s*\((.*)\)\s*$", decl.strip(), flags=re.S)
        if not m:
            continue
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _CTYPE.get(ret.replace("const", "").strip(), ctypes.c_int)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or "[" in a:
                    argtypes.append(ctypes.c_void_p)
                    continue
                if "ctcComputeInfo" in a:
                    argtypes.append(CtcComputeInfo)
                    continue
                toks = a.replace("const", "").split()
                ty = " ".join(toks[:-1]) if len(toks) > 1 else toks[0]
                argtypes.append(_CTYPE[ty] if ty in _CTYPE else ctypes.c_void_p)      # opaque handle typedefs (*_t)
        out[name] = (restype, argtypes)
    return out


def _bind(lib, header):
    decls = parse_header(header)
    for name, (restype, argtypes) in decls.items():
        if os.environ.get("ASLP_B200_ALLOW_MISSING") and not hasattr(lib, name):
            continue                     # bring-up only; the CPU test suite checks every symbol is exported
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    return decls


_cuda = None
_host = None


def cuda_lib():
    """ctypes handle of libaslp_b200.so with argtypes bound from include/aslp_b200.h and include/ctc.h."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(LIB_CUDA_PATH):
            raise RuntimeError("libaslp_b200.so is not built (run python -c 'import __graft_entry__ as g; g.build()'); "
                               "there is no CPU fallback for the aslp-nnet hot path")
        lib = ctypes.CDLL(LIB_CUDA_PATH, mode=ctypes.RTLD_GLOBAL)
        _bind(lib, os.path.join(INCLUDE, "aslp_b200.h"))
        _bind(lib, os.path.join(INCLUDE, "ctc.h"))
        _cuda = lib
    return _cuda


def host_lib():
    """ctypes handle of libaslp_nnet.so (C handle API of the C++ Nnet / IWorker mirror)."""
    global _host
    if _host is None:
        cuda_lib()
        if not os.path.exists(LIB_HOST_PATH):
            raise RuntimeError("libaslp_nnet.so is not built; there is no CPU fallback")
        lib = ctypes.CDLL(LIB_HOST_PATH, mode=ctypes.RTLD_GLOBAL)
        _bind(lib, os.path.join(INCLUDE, "aslp_nnet_c.h"))
        _host = lib
    return _host


def check(status, lib=None):
    if status != 0:
        lib = lib or cuda_lib()
        raise RuntimeError("aslp_b200 call failed (status %d): %s" % (status, lib.aslp_last_error().decode()))


class LstmDir(ctypes.Structure):
    """aslp_lstm_dir_t (include/aslp_b200.h)."""
    _fields_ = [
        ("T", ctypes.c_int), ("S", ctypes.c_int), ("C", ctypes.c_int), ("R", ctypes.c_int),
        ("reverse", ctypes.c_int),
        ("buf", ctypes.c_void_p), ("ldb", ctypes.c_int),
        ("dbuf", ctypes.c_void_p), ("lddb", ctypes.c_int),
        ("w_r", ctypes.c_void_p), ("ldwr", ctypes.c_int),
        ("w_rm", ctypes.c_void_p), ("ldwrm", ctypes.c_int),
        ("peep_i", ctypes.c_void_p), ("peep_f", ctypes.c_void_p), ("peep_o", ctypes.c_void_p),
        ("seq_len_dev", ctypes.c_void_p),
        ("cell_clip", ctypes.c_float),
    ]


class Gru(ctypes.Structure):
    """aslp_gru_t (include/aslp_b200.h)."""
    _fields_ = [
        ("T", ctypes.c_int), ("S", ctypes.c_int), ("H", ctypes.c_int),
        ("buf", ctypes.c_void_p), ("ldb", ctypes.c_int),
        ("dbuf", ctypes.c_void_p), ("lddb", ctypes.c_int),
        ("w_zr_h", ctypes.c_void_p), ("ldwzr", ctypes.c_int),
        ("w_m_g", ctypes.c_void_p), ("ldwmg", ctypes.c_int),
    ]
