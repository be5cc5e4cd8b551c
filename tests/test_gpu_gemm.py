"""GPU parity: aslp_gemm (tcgen05/TMEM/TMA) vs the oracle's restatement of CuMatrixBase::AddMatMat
(src/aslp-cudamatrix/cu-matrix.cc:1027-1062), through the C-ABI."""
import numpy as np
import pytest

from oracle import aslp_oracle as O

pytestmark = pytest.mark.gpu

# tolerance: relative to max|C| of the fp32 oracle.
#   3xTF32 (default): fp32-grade, the north_star's 1e-4 bound (we assert 2e-5)
#   TF32 single pass: stated looser bound 5e-3 (10-bit mantissa operands, truncated)
#   FP32 CUDA cores : 2e-5
TOL = {0: 2e-5, 1: 5e-3, 2: 2e-5, 3: 2e-5}      # 3: fp16 hi / lo planes + three kind::f16 passes (what 0 runs on chunk-sized products)

SHAPES = [
    # (M, N, K)
    (256, 1024, 440),        # cfg1 affine fwd
    (200, 72, 640),          # ragged M, N not a tile multiple (cfg3 output layer)
    (1000, 1280, 40),        # K = 40 (cfg3 layer-1 input projection): K tail zero-fill
    (129, 260, 33),          # everything ragged
    (128, 128, 32),          # exactly one tile, one k-block
]


def run_gemm(M, N, K, ta, tb, prec, alpha=1.0, beta=0.0, bias=False, clip=0.0, seed=0, extra_ld=0):
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync, P
    import torch
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    bv = rng.standard_normal(N).astype(np.float32) if bias else None
    dA, dB, dC = DMat(A, extra_ld=extra_ld), DMat(B, extra_ld=extra_ld), DMat(C0, extra_ld=extra_ld)
    dbias = dvec(np.concatenate([bv, np.zeros(4, np.float32)])) if bias else None
    L = lib()
    wsb = L.aslp_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    ok(L.aslp_gemm(stream(), int(ta), int(tb), M, N, K, alpha, dA.ptr, dA.ld, dB.ptr, dB.ld, beta, dC.ptr, dC.ld,
                   ptr(dbias) if bias else P(0), clip, prec, ptr(ws), wsb))
    sync()
    got = dC.np()
    want = O.gemm(A, B, ta, tb, alpha, beta, C0, bv, clip)
    # double-precision reference for the error scale
    a = (A.T if ta else A).astype(np.float64)
    b = (B.T if tb else B).astype(np.float64)
    exact = alpha * (a @ b) + beta * C0 + (bv[None, :] if bias else 0)
    if clip > 0:
        exact = np.clip(exact, -clip, clip)
    scale = np.abs(exact).max() + 1e-30
    err = np.abs(got - exact).max() / scale
    err_oracle = np.abs(want - exact).max() / scale
    return err, err_oracle, got, dC


@pytest.mark.parametrize("prec", [2, 1, 0, 3], ids=["fp32", "tf32", "x3tf32", "f16x3"])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)], ids=["NT", "NN", "TN", "TT"])
@pytest.mark.parametrize("shape", SHAPES, ids=["s%d" % i for i in range(len(SHAPES))])
def test_gemm_all_layouts(shape, ta, tb, prec):
    M, N, K = shape
    err, err_oracle, _, _ = run_gemm(M, N, K, ta, tb, prec)
    assert err < TOL[prec], (shape, ta, tb, prec, err, err_oracle)


@pytest.mark.parametrize("prec", [0, 1, 3])
def test_gemm_epilogue_alpha_beta_bias_clip(prec):
    err, _, _, _ = run_gemm(300, 200, 96, False, True, prec, alpha=0.5, beta=0.9, bias=True)
    assert err < TOL[prec]
    err, _, got, _ = run_gemm(300, 200, 96, False, True, prec, alpha=1.0, beta=0.9, bias=False, clip=5.0)
    assert err < TOL[prec]
    assert np.abs(got).max() <= 5.0


@pytest.mark.parametrize("prec", [0, 1, 3])
def test_gemm_wgrad_split_k(prec):
    # chunk weight gradient: [4C, D] = DGIFO^T [4C, T*S] * X [T*S, D] with momentum and clip (lc.h:981-1017)
    # long contractions: the bound is the north_star's 1e-4 (fp32-grade paths), the TF32 bound otherwise
    tol = 1e-4 if prec in (0, 3) else TOL[prec]
    err, _, _, _ = run_gemm(1280, 640, 4000, True, False, prec, alpha=1.0, beta=0.9, clip=50.0)
    assert err < tol, err
    err, _, _, _ = run_gemm(320, 40, 4000, True, False, prec, alpha=1.0, beta=0.9)
    assert err < tol, err
    err, _, _, _ = run_gemm(1280, 640, 16000, True, False, prec, alpha=1.0, beta=0.9)      # cfg3 chunk wgrad, full K
    assert err < tol, err


def test_gemm_padded_strides_do_not_leak():
    # ld > cols: the pad columns of C must stay untouched
    err, _, _, dC = run_gemm(130, 70, 64, False, True, 0, extra_ld=8)
    assert err < TOL[0]
    pad = dC.t[:, 70:].cpu().numpy()
    # cols 70..71 belong to the 4-float rounding, 72.. to the extra pad; none may be written
    assert np.all(pad == 0.0)


def test_gemm_empty():
    from tests.gpu_utils import lib, ok, stream, P
    ok(lib().aslp_gemm(stream(), 0, 1, 0, 16, 16, 1.0, P(0), 16, P(0), 16, 0.0, P(0), 16, P(0), 0.0, 0, P(0), 0))


def test_gemm_f16x3_rows_of_very_different_magnitude():
    """the fp16-split path scales every row of op(A) and every column of op(B) by its own power of two: rows 2^-60 .. 2^40 apart must
    all come out fp32-grade RELATIVE TO THEIR OWN magnitude (a single global scale would flush the small rows to zero)"""
    from tests.gpu_utils import DMat, lib, ok, ptr, stream, sync, P
    import torch
    rng = np.random.default_rng(7)
    M, N, K = 300, 200, 256
    A = rng.standard_normal((M, K)).astype(np.float32) * np.exp2(rng.integers(-60, 40, size=(M, 1))).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32) * np.exp2(rng.integers(-30, 30, size=(N, 1))).astype(np.float32)
    dA, dB, dC = DMat(A), DMat(B), DMat(np.zeros((M, N), np.float32))
    L = lib()
    wsb = L.aslp_gemm_workspace_bytes(M, N, K)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    ok(L.aslp_gemm(stream(), 0, 1, M, N, K, 1.0, dA.ptr, dA.ld, dB.ptr, dB.ld, 0.0, dC.ptr, dC.ld, P(0), 0.0, 3, ptr(ws), wsb))
    sync()
    got = dC.np().astype(np.float64)
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(A.astype(np.float64)).max(axis=1)[:, None] * np.abs(B.astype(np.float64)).max(axis=1)[None, :] * np.sqrt(K)
    assert np.all(np.isfinite(got))
    assert (np.abs(got - exact) / scale).max() < 2e-5


@pytest.mark.parametrize("M,N,K,ta,tb", [(256, 1024, 1024, 0, 1), (256, 1024, 1024, 0, 0), (1024, 440, 256, 1, 0), (1000, 512, 1024, 0, 1),
                                          (1000, 1500, 512, 0, 1), (1024, 1024, 1000, 1, 0), (250, 1022, 440, 0, 1),
                                          (4096, 2048, 512, 0, 1), (24, 36, 20, 0, 1)])
def test_gemm_ex_epilogue_matches_the_separate_steps(M, N, K, ta, tb):
    """aslp_gemm_ex: activation of the result, derivative of an activation at its output, SGD apply -- against fp64, for shapes
    that take the split-K reduce pass (folded), a large shape (separate launches behind the product) and a CUDA-core shape."""
    import ctypes
    import torch
    import kaldi_aslp_b200 as KK
    from tests.gpu_utils import lib, ptr, stream, sync
    L = lib()

    class Epi(ctypes.Structure):
        _fields_ = [("act", ctypes.c_int), ("dact_y", ctypes.c_void_p), ("dact_ldy", ctypes.c_int), ("dact_kind", ctypes.c_int),
                    ("update_w", ctypes.c_void_p), ("update_ldw", ctypes.c_int), ("update_lr", ctypes.c_float), ("reduce_in_launch", ctypes.c_int)]
    rng = np.random.default_rng(M + N)
    A = (rng.standard_normal((K, M) if ta else (M, K)) * 0.1).astype(np.float32)
    B = (rng.standard_normal((N, K) if tb else (K, N)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    prod = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    ld = (N + 3) // 4 * 4
    f = ctypes.c_float
    dA, dB, dbias = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), torch.from_numpy(bias).cuda()
    ws = torch.zeros(L.aslp_gemm_workspace_bytes(M, N, K) + 256, dtype=torch.uint8, device="cuda")

    def run(C0, epi, with_bias, beta):
        dC = torch.zeros((M, ld), device="cuda"); dC[:, :N] = torch.from_numpy(C0).cuda()
        rc = L.aslp_gemm_ex(stream(), ta, tb, M, N, K, f(1.0), ptr(dA), A.shape[1], ptr(dB), B.shape[1], f(beta), ptr(dC), ld,
                            ptr(dbias) if with_bias else None, f(0.0), 0, ptr(ws), ctypes.c_size_t(ws.numel() - 256), ctypes.byref(epi))
        assert rc == 0, L.aslp_last_error()
        sync()
        return dC[:, :N].cpu().numpy()
    zero = np.zeros((M, N), np.float32)
    tol = 2e-5
    # every case twice: split-K reduced by the second pass, and inside the launch (reduce_in_launch; a no-op for shapes that
    # do not split or do not fit one wave) -- the two must agree bit for bit (same summation order)
    # (1) forward: sigmoid / tanh / relu of product + bias
    for kind, fn in ((0, lambda x: 1 / (1 + np.exp(-x))), (1, np.tanh), (2, lambda x: np.maximum(x, 0))):
        got = run(zero, Epi(1 + kind, None, 0, 0, None, 0, 0.0, 0), True, 0.0)
        want = fn(prod + bias)
        assert np.abs(got - want).max() < tol * max(1.0, np.abs(want).max()), kind
        assert np.array_equal(run(zero, Epi(1 + kind, None, 0, 0, None, 0, 0.0, 1), True, 0.0), got), kind
    # (2) backward: product times f'(y)
    y = rng.uniform(0.05, 0.95, size=(M, N)).astype(np.float32)
    dy = torch.zeros((M, ld), device="cuda"); dy[:, :N] = torch.from_numpy(y).cuda()
    for kind, fn in ((0, lambda yy: yy * (1 - yy)), (1, lambda yy: 1 - yy * yy), (2, lambda yy: (yy > 0).astype(np.float64))):
        got = run(zero, Epi(0, dy.data_ptr(), ld, kind, None, 0, 0.0, 0), False, 0.0)
        want = fn(y.astype(np.float64)) * prod
        assert np.abs(got - want).max() < tol * max(1.0, np.abs(want).max()), kind
        assert np.array_equal(run(zero, Epi(0, dy.data_ptr(), ld, kind, None, 0, 0.0, 1), False, 0.0), got), kind
    # (3) weight gradient with momentum + SGD apply
    corr0 = rng.standard_normal((M, N)).astype(np.float32)
    W0 = rng.standard_normal((M, N)).astype(np.float32)
    results = []
    for ril in (0, 1, 1):                                  # in-launch twice: the counters re-arm themselves
        dW = torch.zeros((M, ld), device="cuda"); dW[:, :N] = torch.from_numpy(W0).cuda()
        got_corr = run(corr0, Epi(0, None, 0, 0, dW.data_ptr(), ld, 0.01, ril), False, 0.9)
        want_corr = 0.9 * corr0 + prod
        assert np.abs(got_corr - want_corr).max() < tol * np.abs(want_corr).max()
        got_W = dW[:, :N].cpu().numpy()
        assert np.abs(got_W - (W0 - 0.01 * want_corr)).max() < tol * np.abs(W0).max()
        results.append((got_corr, got_W))
    for gc, gw in results[1:]:
        assert np.array_equal(gc, results[0][0]) and np.array_equal(gw, results[0][1])


@pytest.mark.parametrize("rows,cols", [(256, 1024), (1000, 1500), (7, 33), (5000, 512)])
def test_bias_grad_update(rows, cols):
    import ctypes
    import torch
    from tests.gpu_utils import lib, ptr, stream, sync
    L = lib()
    rng = np.random.default_rng(rows)
    ld = (cols + 3) // 4 * 4
    d = rng.standard_normal((rows, cols)).astype(np.float32)
    b0, c0 = rng.standard_normal(cols).astype(np.float32), rng.standard_normal(cols).astype(np.float32)
    dd = torch.zeros((rows, ld), device="cuda"); dd[:, :cols] = torch.from_numpy(d).cuda()
    db, dc = torch.zeros(ld, device="cuda"), torch.zeros(ld, device="cuda")
    db[:cols] = torch.from_numpy(b0).cuda(); dc[:cols] = torch.from_numpy(c0).cuda()
    f = ctypes.c_float
    assert L.aslp_bias_grad_update(stream(), ptr(db), ptr(dc), ptr(dd), ld, rows, cols, f(0.9), f(0.05)) == 0
    sync()
    want_c = 0.9 * c0.astype(np.float64) + d.astype(np.float64).sum(axis=0)
    assert np.abs(dc[:cols].cpu().numpy() - want_c).max() < 1e-5 * max(1.0, np.abs(want_c).max())
    assert np.abs(db[:cols].cpu().numpy() - (b0 - 0.05 * want_c)).max() < 1e-5 * max(1.0, np.abs(want_c).max())
