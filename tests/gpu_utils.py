"""Helpers for the -m gpu parity tests: device buffers come from torch (plumbing), every compute call
goes through the C-ABI of libaslp_b200.so via ctypes."""
import ctypes

import numpy as np
import torch

import kaldi_aslp_b200 as K

P = ctypes.c_void_p


def lib():
    return K.cuda_lib()


def stream():
    return P(torch.cuda.current_stream().cuda_stream)


def ok(status):
    K.check(status)


class DMat:
    """Row-major fp32 device matrix with a padded row stride (multiple of 4 floats)."""

    def __init__(self, arr=None, rows=None, cols=None, extra_ld=0, fill=None):
        if arr is not None:
            arr = np.asarray(arr, np.float32)
            if arr.ndim == 1:
                arr = arr[None, :]
            rows, cols = arr.shape
        self.rows, self.cols = rows, cols
        self.ld = (cols + 3) // 4 * 4 + extra_ld
        self.t = torch.zeros((max(rows, 1), self.ld), dtype=torch.float32, device="cuda")
        if fill is not None:
            self.t.fill_(fill)
        if arr is not None and rows > 0:
            self.t[:rows, :cols] = torch.from_numpy(arr).cuda()

    @property
    def ptr(self):
        return P(self.t.data_ptr())

    def ptr_at(self, r, c=0):
        return P(self.t.data_ptr() + 4 * (r * self.ld + c))

    def np(self):
        return self.t[: self.rows, : self.cols].cpu().numpy()


_KEEP = []   # device tensors handed to the C-ABI by raw pointer must outlive the asynchronous kernels


def dvec(arr, dtype=np.float32):
    a = np.ascontiguousarray(arr, dtype)
    t = torch.from_numpy(a).cuda()
    _KEEP.append(t)
    if len(_KEEP) > 256:
        torch.cuda.synchronize()
        del _KEEP[:128]
    return t


def ptr(t):
    return P(t.data_ptr())


def sync():
    torch.cuda.synchronize()
