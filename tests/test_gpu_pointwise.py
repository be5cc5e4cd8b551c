"""GPU parity for the HBM-bound kernels (activations, softmax, reductions, Xent, Splice, RowConvolution,
BatchNormalization, FSMN, averaging kernels) vs the numpy oracle, through the C-ABI."""
import numpy as np
import pytest

from oracle import aslp_oracle as O

pytestmark = pytest.mark.gpu

KINDS = {"sigmoid": 0, "tanh": 1, "relu": 2}


def close(a, b, tol=1e-5):
    scale = max(np.abs(b).max(), 1e-30)
    return np.abs(a - b).max() / scale < tol


@pytest.mark.parametrize("kind", ["sigmoid", "tanh", "relu"])
@pytest.mark.parametrize("rows,cols", [(256, 1024), (37, 73), (1, 5), (0, 8)])
def test_activations(kind, rows, cols):
    from tests.gpu_utils import DMat, lib, ok, stream, sync
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((rows, cols)) * 4).astype(np.float32)
    e = rng.standard_normal((rows, cols)).astype(np.float32)
    dx, dy, de, dd = DMat(x), DMat(rows=rows, cols=cols), DMat(e), DMat(rows=rows, cols=cols)
    L = lib()
    ok(L.aslp_act_fwd(stream(), KINDS[kind], dy.ptr, dy.ld, dx.ptr, dx.ld, rows, cols))
    src = dx if kind == "relu" else dy
    ok(L.aslp_act_bwd(stream(), KINDS[kind], dd.ptr, dd.ld, src.ptr, src.ld, de.ptr, de.ld, rows, cols))
    sync()
    if rows == 0:
        return
    y = O.act_fwd(kind, x)
    assert close(dy.np(), y, 2e-6)
    assert close(dd.np(), O.act_bwd(kind, x if kind == "relu" else y, e), 5e-6)


@pytest.mark.parametrize("rows,cols", [(256, 1500), (1000, 72), (3, 5)])
def test_softmax_rows(rows, cols):
    from tests.gpu_utils import DMat, lib, ok, stream, sync
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((rows, cols)) * 3).astype(np.float32)
    dx, dy = DMat(x), DMat(rows=rows, cols=cols)
    ok(lib().aslp_softmax_rows(stream(), dy.ptr, dy.ld, dx.ptr, dx.ld, rows, cols))
    sync()
    assert np.abs(dy.np() - O.softmax_rows(x)).max() < 1e-6


def test_axpby_addvec_clamp_transpose_argmax():
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    import torch
    rng = np.random.default_rng(3)
    L = lib()
    a = rng.standard_normal((130, 70)).astype(np.float32)
    b = rng.standard_normal((130, 70)).astype(np.float32)
    v = rng.standard_normal(70).astype(np.float32)
    da, db, dv = DMat(a), DMat(b, extra_ld=4), dvec(np.concatenate([v, np.zeros(4, np.float32)]))
    ok(L.aslp_axpby(stream(), da.ptr, da.ld, db.ptr, db.ld, 130, 70, -0.5, 1.0))
    sync()
    want = a - 0.5 * b
    assert close(da.np(), want)
    ok(L.aslp_add_vec_to_rows(stream(), da.ptr, da.ld, 130, 70, ptr(dv), 2.0, 1.0))
    sync()
    want = want + 2.0 * v[None, :]
    assert close(da.np(), want)
    ok(L.aslp_clamp(stream(), da.ptr, da.ld, 130, 70, -1.0, 1.0))
    sync()
    want = np.clip(want, -1, 1)
    assert close(da.np(), want)
    dt = DMat(rows=70, cols=130)
    ok(L.aslp_transpose(stream(), dt.ptr, dt.ld, da.ptr, da.ld, 130, 70))
    sync()
    assert np.array_equal(dt.np(), want.T)
    idx = torch.zeros(130, dtype=torch.int32, device="cuda")
    ok(L.aslp_row_argmax(stream(), ptr(idx), db.ptr, db.ld, 130, 70))
    sync()
    assert np.array_equal(idx.cpu().numpy(), b.argmax(axis=1))      # bit-exact index bookkeeping


@pytest.mark.parametrize("rows,cols", [(16000, 1280), (256, 1024), (5, 3), (700, 322)])
def test_col_sum_and_col_dot(rows, cols):
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    rng = np.random.default_rng(4)
    a = rng.standard_normal((rows, cols)).astype(np.float32)
    b = rng.standard_normal((rows, cols)).astype(np.float32)
    v0 = rng.standard_normal(cols).astype(np.float32)
    da, db = DMat(a), DMat(b)
    dv = dvec(v0.copy())
    L = lib()
    ok(L.aslp_col_sum(stream(), ptr(dv), da.ptr, da.ld, rows, cols, 1.0, 0.9, 0.0))
    sync()
    want = 0.9 * v0 + a.astype(np.float64).sum(axis=0)
    assert np.abs(dv.cpu().numpy() - want).max() / np.abs(want).max() < 1e-5
    dv = dvec(v0.copy())
    ok(L.aslp_col_dot(stream(), ptr(dv), da.ptr, da.ld, db.ptr, db.ld, rows, cols, 1.0, 0.9, 5.0))
    sync()
    want = np.clip(0.9 * v0 + (a.astype(np.float64) * b).sum(axis=0), -5, 5)
    assert np.abs(dv.cpu().numpy() - want).max() / np.abs(want).max() < 1e-5


def test_l1_and_max_norm():
    from tests.gpu_utils import DMat, lib, ok, stream, sync
    rng = np.random.default_rng(5)
    w = rng.standard_normal((64, 48)).astype(np.float32) * 0.1
    w[3, 4] = 0.0
    g = rng.standard_normal((64, 48)).astype(np.float32)
    dw, dg = DMat(w), DMat(g)
    ok(lib().aslp_regularize_l1(stream(), dw.ptr, dw.ld, dg.ptr, dg.ld, 64, 48, 0.01, 0.05))
    sync()
    w2, g2 = O.regularize_l1(w, g, 0.01, 0.05)
    assert np.array_equal(dw.np(), w2) and np.array_equal(dg.np(), g2)
    ok(lib().aslp_max_norm_rows(stream(), dw.ptr, dw.ld, 64, 48, 0.3))
    sync()
    assert close(dw.np(), O.max_norm_rows(w2, 0.3))


def test_xent_sparse_and_dense():
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    import torch
    rng = np.random.default_rng(6)
    rows, K = 300, 72
    y = O.softmax_rows(rng.standard_normal((rows, K)).astype(np.float32) * 2)
    idx = rng.integers(0, K, rows).astype(np.int32)
    tw = np.ones(rows, np.float32)
    fw = (rng.uniform(size=rows) > 0.2).astype(np.float32)
    tgt = np.zeros((rows, K), np.float32)
    tgt[np.arange(rows), idx] = 1.0
    diff_w, stats_w = O.xent(y, tgt, fw)
    dy, dd = DMat(y), DMat(rows=rows, cols=K)
    stats = torch.zeros(5, dtype=torch.float64, device="cuda")
    ok(lib().aslp_xent_sparse(stream(), dd.ptr, dd.ld, dy.ptr, dy.ld, rows, K, ptr(dvec(idx, np.int32)), ptr(dvec(tw)),
                              ptr(dvec(fw)), ptr(stats)))
    sync()
    assert np.abs(dd.np() - diff_w).max() < 1e-6
    got = stats.cpu().numpy()
    assert np.allclose(got, stats_w, rtol=1e-5, atol=1e-6), (got, stats_w)
    assert got[3] == stats_w[3] and got[4] == stats_w[4]           # frame / correct counters are exact
    # dense soft targets
    tgt2 = rng.uniform(size=(rows, K)).astype(np.float32)
    tgt2 /= tgt2.sum(axis=1, keepdims=True)
    diff_w, stats_w = O.xent(y, tgt2, fw)
    dt = DMat(tgt2)
    stats.zero_()
    ok(lib().aslp_xent_dense(stream(), dd.ptr, dd.ld, dy.ptr, dy.ld, dt.ptr, dt.ld, rows, K, ptr(dvec(fw)), ptr(stats)))
    sync()
    assert np.abs(dd.np() - diff_w).max() < 1e-6
    assert np.allclose(stats.cpu().numpy(), stats_w, rtol=1e-4, atol=1e-5)


def test_splice_fwd_bwd():
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    rng = np.random.default_rng(7)
    rows, D = 53, 40
    offs = list(range(-5, 6))
    x = rng.standard_normal((rows, D)).astype(np.float32)
    dy = rng.standard_normal((rows, D * len(offs))).astype(np.float32)
    dx, do, ddy, ddx = DMat(x), DMat(rows=rows, cols=D * len(offs)), DMat(dy), DMat(rows=rows, cols=D)
    doff = dvec(offs, np.int32)
    ok(lib().aslp_splice_fwd(stream(), do.ptr, do.ld, dx.ptr, dx.ld, rows, D, ptr(doff), len(offs)))
    ok(lib().aslp_splice_bwd(stream(), ddx.ptr, ddx.ld, ddy.ptr, ddy.ld, rows, D, ptr(doff), len(offs)))
    sync()
    assert np.array_equal(do.np(), O.splice_fwd(x, offs))           # a gather: bit-exact
    assert close(ddx.np(), O.splice_bwd(dy, offs, D))


def test_rowconv_fwd_bwd():
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    rng = np.random.default_rng(8)
    T, S, D, Fc = 12, 3, 20, 4
    lens = [12, 7, 1]
    x = rng.standard_normal((T * S, D)).astype(np.float32)
    dy = rng.standard_normal((T * S, D)).astype(np.float32)
    w = rng.standard_normal((D, Fc + 1)).astype(np.float32)
    dx, dw, ddy = DMat(x), DMat(w), DMat(dy)
    dout, ddx, dwd = DMat(rows=T * S, cols=D), DMat(rows=T * S, cols=D), DMat(rows=D, cols=Fc + 1)
    dl = dvec(lens, np.int32)
    L = lib()
    ok(L.aslp_rowconv_fwd(stream(), dout.ptr, dout.ld, dx.ptr, dx.ld, T, S, D, dw.ptr, dw.ld, Fc, ptr(dl)))
    ok(L.aslp_rowconv_bwd(stream(), ddx.ptr, ddx.ld, dwd.ptr, dwd.ld, dx.ptr, dx.ld, ddy.ptr, ddy.ld, T, S, D, dw.ptr, dw.ld, Fc, ptr(dl)))
    sync()
    assert close(dout.np(), O.rowconv_fwd(x, w, S, lens))
    wdx, wwd = O.rowconv_bwd(x, dy, w, S, lens)
    assert close(ddx.np(), wdx) and close(dwd.np(), wwd)


@pytest.mark.parametrize("rows,cols", [(256, 1024), (1000, 72), (33, 7)])
def test_batchnorm(rows, cols):
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    import torch
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.5).astype(np.float32)
    dyv = rng.standard_normal((rows, cols)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cols).astype(np.float32)
    shift = rng.standard_normal(cols).astype(np.float32)
    pad = np.zeros(4, np.float32)
    dx, dout, dxh, ddy, ddx = DMat(x), DMat(rows=rows, cols=cols), DMat(rows=rows, cols=cols), DMat(dyv), DMat(rows=rows, cols=cols)
    dsc, dsh = dvec(np.concatenate([scale, pad])), dvec(np.concatenate([shift, pad]))
    dmean, divs = torch.zeros(cols + 4, device="cuda"), torch.zeros(cols + 4, device="cuda")
    am = torch.zeros(cols, dtype=torch.float64, device="cuda")
    av = torch.zeros(cols, dtype=torch.float64, device="cuda")
    L = lib()
    ok(L.aslp_bn_fwd_train(stream(), dout.ptr, dout.ld, dxh.ptr, dxh.ld, dx.ptr, dx.ld, rows, cols, ptr(dsc), ptr(dsh), 1e-7,
                           ptr(dmean), ptr(divs), ptr(am), ptr(av)))
    sync()
    out, xhat, mean, inv_std, sx, sx2 = O.bn_fwd_train(x, scale, shift)
    assert close(dout.np(), out, 2e-5) and close(dxh.np(), xhat, 2e-5)
    assert np.allclose(am.cpu().numpy(), sx, rtol=1e-9) and np.allclose(av.cpu().numpy(), sx2, rtol=1e-9)   # fp64 running sums
    dscale0 = rng.standard_normal(cols).astype(np.float32)
    dshift0 = rng.standard_normal(cols).astype(np.float32)
    dds, ddsh = dvec(dscale0.copy()), dvec(dshift0.copy())
    ok(L.aslp_bn_bwd(stream(), ddx.ptr, ddx.ld, dx.ptr, dx.ld, dxh.ptr, dxh.ld, ddy.ptr, ddy.ld, rows, cols, ptr(dsc), ptr(dmean),
                     ptr(divs), 0.9, ptr(dds), ptr(ddsh)))
    sync()
    wdx, wds, wdsh = O.bn_bwd(x, xhat, dyv, scale, mean, inv_std, 0.9, dscale0, dshift0)
    assert close(ddx.np(), wdx, 1e-4), np.abs(ddx.np() - wdx).max()
    assert close(dds.cpu().numpy(), wds, 1e-4) and close(ddsh.cpu().numpy(), wdsh, 1e-4)
    # eval path
    ok(L.aslp_bn_fwd_eval(stream(), dout.ptr, dout.ld, dx.ptr, dx.ld, rows, cols, ptr(dsc), ptr(dsh), ptr(dmean), ptr(divs)))
    sync()
    assert close(dout.np(), O.bn_fwd_eval(x, scale, shift, mean, inv_std), 2e-5)


@pytest.mark.parametrize("T,D,P,Fu", [(1000, 512, 20, 20), (37, 70, 3, 5), (5, 8, 20, 20), (200, 128, 0, 7), (130, 36, 30, 30)])
def test_fsmn(T, D, P, Fu):
    from tests.gpu_utils import DMat, lib, ok, stream, sync
    rng = np.random.default_rng(10)
    C = P + Fu + 1
    x = rng.standard_normal((T, D)).astype(np.float32)
    dyv = rng.standard_normal((T, D)).astype(np.float32)
    coef = (rng.standard_normal((C, D)) * 0.2).astype(np.float32)
    dx, dc, ddy = DMat(x), DMat(coef), DMat(dyv)
    dout, ddx, dcorr = DMat(rows=T, cols=D), DMat(rows=T, cols=D), DMat(rows=C, cols=D)
    L = lib()
    ok(L.aslp_fsmn_fwd(stream(), dout.ptr, dout.ld, dx.ptr, dx.ld, T, D, dc.ptr, dc.ld, P, Fu))
    ok(L.aslp_fsmn_bwd(stream(), ddx.ptr, ddx.ld, ddy.ptr, ddy.ld, T, D, dc.ptr, dc.ld, P, Fu))
    ok(L.aslp_fsmn_coef_grad(stream(), dcorr.ptr, dcorr.ld, dx.ptr, dx.ld, ddy.ptr, ddy.ld, T, D, P, Fu, 0.0))
    sync()
    assert close(dout.np(), O.fsmn_fwd(x, coef, P, Fu), 1e-5)
    wdx, wcorr = O.fsmn_bwd(x, dyv, coef, P, Fu)
    assert close(ddx.np(), wdx, 1e-5)
    assert close(dcorr.np(), wcorr, 1e-4)
    # linearity (size-independent property): filter(a*x) == a*filter(x)
    dx2 = DMat(2.0 * x)
    dout2 = DMat(rows=T, cols=D)
    ok(L.aslp_fsmn_fwd(stream(), dout2.ptr, dout2.ld, dx2.ptr, dx2.ld, T, D, dc.ptr, dc.ld, P, Fu))
    sync()
    assert np.array_equal(dout2.np(), 2.0 * dout.np())


def test_sync_kernels():
    from tests.gpu_utils import dvec, lib, ok, ptr, stream, sync
    rng = np.random.default_rng(11)
    n = 100003
    L = lib()
    w = [rng.standard_normal(n).astype(np.float32) for _ in range(4)]
    w_prev = rng.standard_normal(n).astype(np.float32)
    delta_prev = rng.standard_normal(n).astype(np.float32) * 0.1
    # BMUF with the allreduce restated as a host sum of per-rank (w - w_prev)
    g = np.zeros(n, np.float32)
    for wi in w:
        dg = dvec(np.zeros(n, np.float32))
        ok(L.aslp_sync_diff(stream(), ptr(dg), ptr(dvec(wi)), ptr(dvec(w_prev)), n))
        sync()
        g += dg.cpu().numpy()
    dw, dwp, ddp = dvec(w[0].copy()), dvec(w_prev.copy()), dvec(delta_prev.copy())
    ok(L.aslp_sync_bmuf_apply(stream(), ptr(dw), ptr(dwp), ptr(ddp), ptr(dvec(g)), n, 0.75, 1.0))
    sync()
    ww, wwp, wdp = O.bmuf_sync(w, w_prev, delta_prev, 0.75, 1.0)
    assert np.allclose(dw.cpu().numpy(), ww, rtol=1e-5, atol=1e-6)
    assert np.array_equal(dw.cpu().numpy(), dwp.cpu().numpy())
    assert np.allclose(ddp.cpu().numpy(), wdp, rtol=1e-5, atol=1e-6)
    # BSP scale
    ds = dvec(np.zeros(n, np.float32))
    ok(L.aslp_sync_scale(stream(), ptr(ds), ptr(dvec(w[1])), n, 0.25))
    sync()
    assert np.array_equal(ds.cpu().numpy(), w[1] * np.float32(0.25))
    # SOD optimizers
    for oi, name in enumerate(["sgd", "momentum", "adagrad", "rmsprop", "adadelta", "adam"]):
        s1 = np.abs(rng.standard_normal(n)).astype(np.float32) * 0.1
        s2 = np.abs(rng.standard_normal(n)).astype(np.float32) * 0.1
        gg = rng.standard_normal(n).astype(np.float32)
        p1 = {"momentum": 0.9, "adadelta": 0.95, "adam": 0.9}.get(name, 0.0)
        dw, ds1, ds2 = dvec(w[2].copy()), dvec(s1.copy()), dvec(s2.copy())
        ok(L.aslp_sync_sod_apply(stream(), oi, ptr(dw), ptr(dvec(gg)), ptr(ds1), ptr(ds2), n, 0.01, p1, 0.999, 1e-8, 3))
        sync()
        ww, ws1, ws2 = O.sod_optimize(name, w[2], gg, s1, s2, 0.01, p1, 0.999, 3)
        assert np.allclose(dw.cpu().numpy(), ww, rtol=2e-5, atol=1e-6), name


@pytest.mark.parametrize("rows,cols,nsrc", [(37, 40, 50), (256, 441, 300), (5, 3, 9)])
def test_copy_rows_is_an_exact_gather(rows, cols, nsrc):
    """CuMatrixBase::CopyRows / cu::Randomize (the frame shuffle of MatrixRandomizer): dst[r] = src[idx[r]], idx < 0 -> zeros."""
    from tests.gpu_utils import DMat, P, dvec, lib, ok, stream, sync
    rng = np.random.default_rng(rows + cols)
    src = rng.standard_normal((nsrc, cols)).astype(np.float32)
    idx = rng.integers(0, nsrc, size=rows).astype(np.int32)
    idx[rows // 2] = -1
    S, D = DMat(src), DMat(rows=rows, cols=cols, fill=3.0)
    ok(lib().aslp_copy_rows(stream(), D.ptr, D.ld, S.ptr, S.ld, P(dvec(idx, np.int32).data_ptr()), rows, cols))
    sync()
    want = src[np.maximum(idx, 0)]
    want[idx < 0] = 0
    assert np.array_equal(D.np(), want)               # pure data movement: bit-exact


@pytest.mark.parametrize("rows,cols,apply_log,blank,prior", [(300, 72, 1, 0.5, True), (64, 1500, 1, 0.0, False), (17, 7, 0, 0.0, True)])
def test_posterior_finalize(rows, cols, apply_log, blank, prior):
    """Forwarder tail in one pass (aslp-nnet-forward.cc:184-207): log(x + 1e-20), blank column shift, prior subtraction and
    the min / max / non-finite statistics the warnings use."""
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    import torch
    rng = np.random.default_rng(21)
    x = O.softmax_rows(rng.standard_normal((rows, cols)).astype(np.float32) * 3)
    lp = np.log(rng.uniform(0.01, 1.0, cols)).astype(np.float32)
    dx = DMat(x)
    stats = torch.zeros(8, dtype=torch.float32, device="cuda")
    ok(lib().aslp_posterior_finalize(stream(), dx.ptr, dx.ld, rows, cols, apply_log, 1e-20, blank, ptr(dvec(lp)) if prior else None, 0.8,
                                     ptr(stats)))
    sync()
    want = x.copy()
    if apply_log:
        want = np.log(want + np.float32(1e-20)).astype(np.float32)
    if blank > 0:
        want[:, 0] -= np.float32(blank)
    mid = want.copy()
    if prior:
        want = want + np.float32(-0.8) * lp[None, :]
    assert np.abs(dx.np() - want).max() < 2e-6 * max(1.0, np.abs(want).max())
    s = stats.cpu().numpy()
    assert s[0] == x.min() and s[1] == x.max()
    assert abs(s[2] - mid.min()) < 1e-5 * max(1.0, abs(mid.min())) and abs(s[3] - mid.max()) < 1e-5
    assert s[4] == 0


@pytest.mark.parametrize("rows,ns,stride,pd,step,nf", [(256, 11, 40, 9, 1, 128), (37, 3, 11, 4, 1, 6), (50, 2, 12, 4, 2, 5), (9, 1, 8, 8, 1, 3),
                                                         (8000, 11, 40, 9, 1, 8), (300, 20, 64, 8, 1, 4)])
def test_conv_gather_gemm_scatter(rows, ns, stride, pd, step, nf):
    """ConvolutionalComponent device side (nnet-convolutional-component.h:263-421): im2col into [frames*P, filter_dim], ONE GEMM
    into the [frames, P*num_filters] output, and the inverse gather-sum; the first shape is the recipes' 40 x 11 input."""
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync
    rng = np.random.default_rng(rows + nf)
    in_dim = ns * stride
    npatch, fd = 1 + (stride - pd) // step, ns * pd
    x = rng.standard_normal((rows, in_dim)).astype(np.float32)
    filt = (rng.standard_normal((nf, fd)) * 0.2).astype(np.float32)
    bias = rng.standard_normal(nf).astype(np.float32)
    want_out, want_vec = O.conv_fwd(x, filt, bias, pd, step, stride)
    dx, dpat, dfilt = DMat(x), DMat(rows=rows * npatch, cols=fd), DMat(filt)
    ok(lib().aslp_conv_gather_patches(stream(), dpat.ptr, dpat.ld, dx.ptr, dx.ld, rows, npatch, ns, pd, step, stride))
    sync()
    assert np.array_equal(dpat.np().reshape(rows, npatch * fd), want_vec)             # a pure gather: bit-exact
    import torch
    out = torch.zeros((rows * npatch, nf), dtype=torch.float32, device="cuda")       # dense view of [rows, P*nf]
    db = dvec(bias)
    ok(lib().aslp_gemm(stream(), 0, 1, rows * npatch, nf, fd, 1.0, dpat.ptr, dpat.ld, dfilt.ptr, dfilt.ld, 0.0, ptr(out), nf, ptr(db), 0.0, 0, None, 0))
    sync()
    got_out = out.cpu().numpy().reshape(rows, npatch * nf)
    assert np.abs(got_out - want_out).max() < 1e-4 * max(1.0, np.abs(want_out).max())
    od = (rng.standard_normal((rows, npatch * nf)) * 0.1).astype(np.float32)
    want_in, want_pd = O.conv_bwd(od, filt, in_dim, pd, step, stride)
    dpd = DMat(want_pd.reshape(rows * npatch, fd))
    din = DMat(rows=rows, cols=in_dim, fill=7.0)                                       # overwritten, not accumulated
    ok(lib().aslp_conv_scatter_patch_diffs(stream(), din.ptr, din.ld, dpd.ptr, dpd.ld, rows, npatch, ns, pd, step, stride))
    sync()
    assert np.array_equal(din.np(), want_in)                                           # same summation order: bit-exact


@pytest.mark.parametrize("rows,patches,size,step,ps", [(256, 32, 4, 4, 128), (40, 8, 4, 2, 6), (13, 7, 3, 2, 5), (5, 4, 4, 1, 3), (300, 9, 3, 2, 16)])
def test_maxpool_fwd_bwd(rows, patches, size, step, ps):
    """MaxPoolingComponent (nnet-max-pooling-component.h:100-156), overlapping pools and ties included."""
    from tests.gpu_utils import DMat, lib, ok, stream, sync
    rng = np.random.default_rng(rows * 7 + ps)
    pools = 1 + (patches - size) // step
    x = rng.standard_normal((rows, patches * ps)).astype(np.float32)
    x[::3, :] = np.round(x[::3, :])                                                    # ties: several members equal the maximum
    want = O.maxpool_fwd(x, size, step, ps)
    dx, dout = DMat(x), DMat(rows=rows, cols=pools * ps)
    ok(lib().aslp_maxpool_fwd(stream(), dout.ptr, dout.ld, dx.ptr, dx.ld, rows, pools, size, step, ps))
    sync()
    assert np.array_equal(dout.np(), want)
    od = rng.standard_normal((rows, pools * ps)).astype(np.float32)
    want_in = O.maxpool_bwd(x, want, od, size, step, ps)
    dod, din = DMat(od), DMat(rows=rows, cols=patches * ps, fill=3.0)
    ok(lib().aslp_maxpool_bwd(stream(), din.ptr, din.ld, dx.ptr, dx.ld, dout.ptr, dout.ld, dod.ptr, dod.ld, rows, patches, pools, size, step, ps))
    sync()
    assert np.array_equal(din.np(), want_in)
