"""CPU-side checks of the drop-in boundary (no compute, no GPU needed):
* both in-tree shared objects load and export every symbol include/*.h declares;
* the product fails LOUDLY without a CUDA device (no CPU fallback anywhere below the host layer);
* nothing in the product tree imports, links or mentions oracle/ (the oracle is test infrastructure only);
* every C-ABI entry point group cites the reference interface it replaces."""
import ctypes
import os
import re
import subprocess

import pytest

import kaldi_aslp_b200 as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = {"aslp_b200.h": K.LIB_CUDA_PATH, "ctc.h": K.LIB_CUDA_PATH, "aslp_nnet_c.h": K.LIB_HOST_PATH}


def exported(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


@pytest.mark.parametrize("header", sorted(HEADERS))
def test_every_declared_symbol_is_exported(header):
    lib = HEADERS[header]
    assert os.path.exists(lib), "%s not built: run __graft_entry__.build()" % lib
    decls = K.parse_header(os.path.join(K.INCLUDE, header))
    assert len(decls) >= (3 if header == "ctc.h" else 40), (header, len(decls))
    syms = exported(lib)
    missing = sorted(n for n in decls if n not in syms)
    assert not missing, missing


def test_libraries_load_and_bind_with_ctypes():
    cu, host = K.cuda_lib(), K.host_lib()
    assert cu.aslp_last_error is not None and host.aslp_nnet_last_error is not None
    n = ctypes.c_int(-1)
    rc = cu.aslp_device_count(ctypes.byref(n))
    assert (rc == 0 and n.value >= 0) or rc != 0        # never crashes, with or without a driver


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="needs a machine WITHOUT a CUDA device")
def test_product_fails_loudly_without_a_device(tmp_path):
    cu = K.cuda_lib()
    p = ctypes.c_void_p()
    rc = cu.aslp_malloc(ctypes.byref(p), 1024)
    assert rc != 0 and len(cu.aslp_last_error()) > 0
    # the host layer: reading a model allocates device parameters -> must fail with a message, not fall back
    host = K.host_lib()
    h = ctypes.c_void_p()
    model = os.path.join(ROOT, "tests", "golden", "dnn_xent", "model.bin")
    rc = host.aslp_nnet_read(model.encode(), ctypes.byref(h))
    assert rc != 0
    msg = host.aslp_nnet_last_error().decode()
    assert msg, "failure without a diagnostic"
    # warp-ctc entry point: CPU location is refused outright
    info = K.CtcComputeInfo(0, None)                    # CTC_CPU
    one = (ctypes.c_int * 1)(1)
    f = (ctypes.c_float * 4)()
    rc = cu.compute_ctc_loss(f, f, one, one, one, 4, 1, f, f, info)
    assert rc != 0


def test_product_tree_never_touches_the_oracle():
    bad = []
    for base in (os.path.join(ROOT, "kaldi-aslp_b200"), os.path.join(ROOT, "include")):
        for dp, dn, fn in os.walk(base):
            dn[:] = [d for d in dn if d not in ("build", "__pycache__")]
            for f in fn:
                if not f.endswith((".py", ".cc", ".h", ".cu", ".cuh")):
                    continue
                txt = open(os.path.join(dp, f), errors="replace").read()
                # an import, an include, a path or a link name (plain prose such as an error message is fine)
                if re.search(r"(from|import)\s+oracle|oracle/|oracle\.|ctc_oracle|aslp_oracle|aslp_ref|ref_driver", txt) or "/root/reference" in txt:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
    # and the shared objects do not link anything from oracle/_ref
    for lib in (K.LIB_CUDA_PATH, K.LIB_HOST_PATH):
        ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "aslp_ref" not in ldd and "oracle" not in ldd, ldd


def test_headers_cite_the_reference_interfaces_they_replace():
    for header, minimum in (("aslp_b200.h", 30), ("aslp_nnet_c.h", 10), ("ctc.h", 1)):
        txt = open(os.path.join(K.INCLUDE, header)).read()
        cites = re.findall(r"[\w\-/]+\.(?:h|cc|cu|cpp):\d+", txt)
        assert len(cites) >= minimum, (header, len(cites))
