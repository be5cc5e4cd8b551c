"""GEMM timing at the shapes of the cfg3 / cfg1 training steps through the C-ABI (aslp_gemm), for A/B runs of kernel variants:
  ASLP_B200_CUDA_LIB=kaldi-aslp_b200/libaslp_b200_<variant>.so python tools/gemm_bench.py
CUDA events around `reps` back-to-back launches after warm-up; operands rotate over enough buffers to exceed L2."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import kaldi_aslp_b200 as K  # noqa: E402

P = ctypes.c_void_p
L = K.cuda_lib()
# (name, trans_a, trans_b, M, N, K): fwd x W^T, bwd-data dY W, wgrad dY^T X at the cfg3 layer sizes; cfg1 minibatch shapes
SHAPES = [("cfg3_fwd_xWx", 0, 1, 16000, 1280, 640), ("cfg3_bwd_data", 0, 0, 16000, 640, 1280), ("cfg3_wgrad", 1, 0, 1280, 640, 16000),
          ("cfg3_r_proj", 0, 1, 16000, 320, 320), ("cfg1_fwd", 0, 1, 256, 1024, 1024), ("cfg1_wgrad", 1, 0, 1024, 1024, 256),
          ("square_4096", 0, 1, 4096, 4096, 4096)]


def main():
    st = P(torch.cuda.current_stream().cuda_stream)
    out = {"lib": os.environ.get("ASLP_B200_CUDA_LIB", "product")}
    for name, ta, tb, M, N, Kd in SHAPES:
        for prec, pname in ((0, "3xtf32"), (3, "f16x3"), (1, "tf32")):
            nset = max(2, int(3e8 // ((M * Kd + N * Kd + M * N) * 4)) + 1)
            nset = min(nset, 24)
            As = [torch.randn((Kd, M) if ta else (M, Kd), device="cuda") for _ in range(nset)]
            Bs = [torch.randn((N, Kd) if tb else (Kd, N), device="cuda") for _ in range(nset)]
            Cs = [torch.zeros(M, N, device="cuda") for _ in range(nset)]
            wsb = L.aslp_gemm_workspace_bytes(M, N, Kd)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")

            def launch(i):
                a, b, c = As[i % nset], Bs[i % nset], Cs[i % nset]
                K.check(L.aslp_gemm(st, ta, tb, M, N, Kd, 1.0, P(a.data_ptr()), a.shape[1], P(b.data_ptr()), b.shape[1], 0.0, P(c.data_ptr()), N,
                                    None, 0.0, prec, P(ws.data_ptr()), wsb))
            for i in range(5):
                launch(i)
            torch.cuda.synchronize()
            reps = 40
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                launch(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out["%s_%s" % (name, pname)] = {"ms": round(ms, 5), "useful_TFLOPs": round(2.0 * M * N * Kd / ms / 1e9, 1)}
            del As, Bs, Cs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
