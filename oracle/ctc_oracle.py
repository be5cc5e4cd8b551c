"""ctypes access to (a) oracle/ctc_oracle.c, the plain-C restatement of the warp-ctc CPU algorithm, and
(b) the UNMODIFIED reference's compute_ctc_loss in oracle/_ref/libaslp_ref.so when it has been built.
TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libctc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libaslp_ref.so")


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(HERE, "ctc_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, src, "-lm"])
    return SO


def _prep(acts, labels, input_lengths):
    acts = np.ascontiguousarray(acts, np.float32)          # [maxT, mb, K]
    maxT, mb, K = acts.shape
    flat = np.ascontiguousarray(np.concatenate([np.asarray(l, np.int32) for l in labels]) if len(labels) else np.zeros(0, np.int32))
    if flat.size == 0:
        flat = np.zeros(1, np.int32)
    llen = np.ascontiguousarray([len(l) for l in labels], np.int32)
    ilen = np.ascontiguousarray(input_lengths, np.int32)
    return acts, flat, llen, ilen, maxT, mb, K


def cost_and_grad(acts, labels, input_lengths):
    """acts [maxT, mb, K] unnormalised; labels list of int lists; -> (costs [mb], grads [maxT, mb, K])."""
    lib = ctypes.CDLL(build())
    acts, flat, llen, ilen, maxT, mb, K = _prep(acts, labels, input_lengths)
    grads = np.zeros_like(acts)
    costs = np.zeros(mb, np.float32)
    P = ctypes.c_void_p
    lib.ctc_oracle_cost_and_grad.argtypes = [P, P, P, P, P, ctypes.c_int, ctypes.c_int, P]
    rc = lib.ctc_oracle_cost_and_grad(acts.ctypes.data, grads.ctypes.data, flat.ctypes.data, llen.ctypes.data,
                                      ilen.ctypes.data, K, mb, costs.ctypes.data)
    assert rc == 0
    return costs, grads


class _Info(ctypes.Structure):
    _fields_ = [("loc", ctypes.c_int), ("num_threads", ctypes.c_uint), ("_pad", ctypes.c_uint)]


def have_ref():
    return os.path.exists(REF_SO)


def ref_cost_and_grad(acts, labels, input_lengths, num_threads=1):
    """The reference's own CPU compute_ctc_loss (src/warp-ctc/src/ctc_entrypoint.cpp:35-88), CTC_CPU."""
    lib = ctypes.CDLL(REF_SO, mode=os.RTLD_LAZY)
    acts, flat, llen, ilen, maxT, mb, K = _prep(acts, labels, input_lengths)
    grads = np.zeros_like(acts)
    costs = np.zeros(mb, np.float32)
    P = ctypes.c_void_p
    info = _Info(0, num_threads, 0)
    size = ctypes.c_size_t(0)
    lib.get_workspace_size.argtypes = [P, P, ctypes.c_int, ctypes.c_int, _Info, P]
    lib.compute_ctc_loss.argtypes = [P, P, P, P, P, ctypes.c_int, ctypes.c_int, P, P, _Info]
    assert lib.get_workspace_size(llen.ctypes.data, ilen.ctypes.data, K, mb, info, ctypes.addressof(size)) == 0
    ws = np.zeros(size.value // 4 + 16, np.float32)
    rc = lib.compute_ctc_loss(acts.ctypes.data, grads.ctypes.data, flat.ctypes.data, llen.ctypes.data, ilen.ctypes.data,
                              K, mb, costs.ctypes.data, ws.ctypes.data, info)
    assert rc == 0
    return costs, grads
