"""Phase stamps of the in-launch split-K reduction (profiles/r02_gemm_in_launch_reduce.txt).
Needs a library built with the stamps:  tools/build_variant.sh dbg gemm.cu -DASLP_GEMM_DEBUG_TIMES
then  ASLP_B200_CUDA_LIB=$PWD/kaldi-aslp_b200/libaslp_b200_dbg.so python tools/gemm_in_launch_probe.py"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, ".")
from tests.gpu_utils import lib, ptr, stream, sync
L = lib()
class Epi(ctypes.Structure):
    _fields_ = [("act", ctypes.c_int), ("dact_y", ctypes.c_void_p), ("dact_ldy", ctypes.c_int), ("dact_kind", ctypes.c_int),
                ("update_w", ctypes.c_void_p), ("update_ldw", ctypes.c_int), ("update_lr", ctypes.c_float), ("reduce_in_launch", ctypes.c_int)]
f = ctypes.c_float
print("clusters that fit at once, by size:", {cs: L.aslp_gemm_cluster_fit(cs) for cs in range(2, 9)})
for (M, N, K, ta, tb) in [(256, 1024, 1024, 0, 1), (1024, 1024, 256, 1, 0)]:
    A = torch.randn((K, M) if ta else (M, K), device="cuda") * 0.1
    B = torch.randn((N, K) if tb else (K, N), device="cuda") * 0.1
    C = torch.zeros((M, N), device="cuda")
    ws = torch.zeros(L.aslp_gemm_workspace_bytes(M, N, K) + 256, dtype=torch.uint8, device="cuda")
    for ril in (0, 1):
        epi = Epi(0, None, 0, 0, None, 0, 0.0, ril)
        def call():
            rc = L.aslp_gemm_ex(stream(), ta, tb, M, N, K, f(1.0), ptr(A), A.shape[1], ptr(B), B.shape[1], f(0.0), ptr(C), N,
                                None, f(0.0), 0, ptr(ws), ctypes.c_size_t(ws.numel() - 256), ctypes.byref(epi))
            assert rc == 0
        for _ in range(5): call()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): call()
        e1.record(); sync()
        print(M, N, K, "ril", ril, "us/call", round(e0.elapsed_time(e1) * 1000 / 50, 2))
        if ril:
            call(); sync()
            buf = (ctypes.c_ulonglong * (148 * 8))()
            L.aslp_gemm_debug_times(buf)
            t = np.array(buf[:], dtype=np.int64).reshape(148, 8)
            t = t[t[:, 0] > 0]
            base = t[:, 0].min()
            print(" kernel entry -> set-up done (barriers, TMEM, block sync), ns, median / max:", np.median(t[:, 0] - t[:, 5]), (t[:, 0] - t[:, 5]).max())
            d = t[:, [0, 4, 1, 2, 3]] - base
            print(" ctas", len(t), "start/tmem_full/partial_done/spin_done/reduce_done ns: median", np.median(d, 0), "max", d.max(0), "min", d.min(0))
