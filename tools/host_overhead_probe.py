"""Host-side cost of ENQUEUEING one call through the C-ABI (no synchronisation inside the loop): where the launch-bound
configs (cfg1 / cfg4) spend their host time.  Prints microseconds per call."""
import ctypes, json, sys, time
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import kaldi_aslp_b200 as K
L = K.cuda_lib()
P = ctypes.c_void_p
st = P(torch.cuda.current_stream().cuda_stream)
a = torch.randn(256, 1024, device="cuda"); b = torch.randn(1024, 1024, device="cuda"); c = torch.zeros(256, 1024, device="cuda")
w = torch.zeros(1024, 1024, device="cuda"); v = torch.zeros(1024, device="cuda")
ws = torch.zeros(64 << 20, dtype=torch.uint8, device="cuda")
f = ctypes.c_float
def t(name, fn, n=2000):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(json.dumps({"call": name, "host_us_per_call": round((t1 - t0) / n * 1e6, 2), "us_per_call_incl_gpu": round((t2 - t0) / n * 1e6, 2)}), flush=True)
t("axpby 256x1024", lambda: L.aslp_axpby(st, P(c.data_ptr()), 1024, P(a.data_ptr()), 1024, 256, 1024, f(1.0), f(0.5)))
t("act_fwd sigmoid 256x1024", lambda: L.aslp_act_fwd(st, 0, P(c.data_ptr()), 1024, P(a.data_ptr()), 1024, 256, 1024))
t("col_sum 256x1024", lambda: L.aslp_col_sum(st, P(v.data_ptr()), P(a.data_ptr()), 1024, 256, 1024, f(1.0), f(0.9), f(0.0)))
t("gemm fwd 256x1024x1024 (NT)", lambda: L.aslp_gemm(st, 0, 1, 256, 1024, 1024, f(1.0), P(a.data_ptr()), 1024, P(b.data_ptr()), 1024, f(0.0), P(c.data_ptr()), 1024, P(v.data_ptr()), f(0.0), 0, None, ctypes.c_size_t(0)))
t("gemm bwd 256x1024x1024 (NN)", lambda: L.aslp_gemm(st, 0, 0, 256, 1024, 1024, f(1.0), P(a.data_ptr()), 1024, P(b.data_ptr()), 1024, f(0.0), P(c.data_ptr()), 1024, None, f(0.0), 0, None, ctypes.c_size_t(0)))
t("gemm wgrad 1024x1024x256 (TN) + ws", lambda: L.aslp_gemm(st, 1, 0, 1024, 1024, 256, f(1.0), P(a.data_ptr()), 1024, P(c.data_ptr()), 1024, f(0.9), P(w.data_ptr()), 1024, None, f(0.0), 0, P(ws.data_ptr()), ctypes.c_size_t(64 << 20)))
t("python no-op ctypes call (aslp_num... baseline)", lambda: L.aslp_last_error())
