"""CPU restatement (numpy, fp32) of the reference's algorithms for the aslp-nnet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under kaldi-aslp_b200/ may import this module; it is the
checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.

Every function cites the reference file:line it restates (paths relative to /root/reference/).
Pinning: tests/test_oracle_vs_golden.py checks these functions against fixtures produced by the
UNMODIFIED reference CPU build (oracle/_ref, see oracle/Makefile + oracle/make_golden.py) and the
CTC restatement (oracle/ctc_oracle.c) against warp-ctc's own known-answer tests
(src/warp-ctc/tests/test_cpu.cpp:12-242).
"""
import numpy as np

F = np.float32


# ----------------------------------------------------------------------------- scalar functions
def sigmoid(x):
    """VectorBase::Sigmoid, src/matrix/kaldi-vector.cc:923-936 (overflow-safe split)."""
    x = np.asarray(x, F)
    out = np.empty_like(x)
    pos = x > 0
    out[pos] = F(1) / (F(1) + np.exp(-x[pos], dtype=F))
    ex = np.exp(x[~pos], dtype=F)
    out[~pos] = ex / (ex + F(1))
    return out


def tanh(x):
    """VectorBase::Tanh, src/matrix/kaldi-vector.cc:885-898."""
    x = np.asarray(x, F)
    out = np.empty_like(x)
    pos = x > 0
    ie = np.exp(-x[pos], dtype=F)
    out[pos] = F(-1) + F(2) / (F(1) + ie * ie)
    ie = np.exp(x[~pos], dtype=F)
    out[~pos] = F(1) - F(2) / (F(1) + ie * ie)
    return out


def softmax_rows(x):
    """ApplySoftMaxPerRow -> VectorBase::ApplySoftMax, src/matrix/kaldi-vector.cc:852-859."""
    x = np.asarray(x, F)
    e = np.exp(x - x.max(axis=1, keepdims=True), dtype=F)
    return (e / e.sum(axis=1, keepdims=True, dtype=F)).astype(F)


def act_fwd(kind, x):
    """Sigmoid/Tanh/ReLU::PropagateFnc, src/aslp-nnet/nnet-activation.h:164-167,189-192,287-291."""
    return {"sigmoid": sigmoid, "tanh": tanh, "relu": lambda v: np.maximum(np.asarray(v, F), F(0))}[kind](x)


def act_bwd(kind, y_or_x, e):
    """BackpropagateFnc: y(1-y)e, (1-y^2)e on the OUTPUT; ReLU Heaviside(x)*e on the INPUT
    (nnet-activation.h:169-173,194-198,293-297; cu-kernels.cu:1814,1846,1346)."""
    y, e = np.asarray(y_or_x, F), np.asarray(e, F)
    if kind == "sigmoid":
        return (y * (F(1) - y) * e).astype(F)
    if kind == "tanh":
        return ((F(1) - y * y) * e).astype(F)
    return np.where(y > 0, e, F(0)).astype(F)


# ----------------------------------------------------------------------------- affine
def gemm(A, B, trans_a=False, trans_b=False, alpha=1.0, beta=0.0, C=None, bias=None, clip=0.0):
    """CuMatrixBase::AddMatMat, src/aslp-cudamatrix/cu-matrix.cc:1027-1062 (CPU: cblas_sgemm)."""
    a = np.asarray(A, F).T if trans_a else np.asarray(A, F)
    b = np.asarray(B, F).T if trans_b else np.asarray(B, F)
    out = F(alpha) * (a @ b)
    if C is not None and beta != 0.0:
        out = out + F(beta) * np.asarray(C, F)
    if bias is not None:
        out = out + np.asarray(bias, F)[None, :]
    if clip > 0:
        out = np.clip(out, -clip, clip)
    return out.astype(F)


def affine_step(W, b, Wc, bc, x, dy, lr, mmt, l2=0.0, l1=0.0, lr_coef=1.0, bias_lr_coef=1.0, max_norm=0.0):
    """AffineTransform Propagate / Backpropagate / Update, src/aslp-nnet/nnet-affine-transform.h:186-245.
    Returns (y, dx, W', b', Wc', bc')."""
    W, b, Wc, bc, x, dy = (np.asarray(v, F) for v in (W, b, Wc, bc, x, dy))
    y = (b[None, :] + x @ W.T).astype(F)                           # :186-191
    dx = (dy @ W).astype(F)                                        # :193-197
    n = x.shape[0]
    Wc = (F(mmt) * Wc + dy.T @ x).astype(F)                        # :210
    bc = (F(mmt) * bc + dy.sum(axis=0, dtype=F)).astype(F)         # :211
    lrw, lrb = F(lr * lr_coef), F(lr * bias_lr_coef)
    W = W.copy()
    if l2 != 0.0:
        W = (W + F(-lrw * l2 * n) * W).astype(F)                   # :213-215
    if l1 != 0.0:
        W, Wc = regularize_l1(W, Wc, F(lrw * l1 * n), lrw)         # :217-219
    W = (W - lrw * Wc).astype(F)                                   # :229
    b = (b - lrb * bc).astype(F)                                   # :230
    if max_norm > 0.0:
        W = max_norm_rows(W, max_norm)                             # :232-243
    return y, dx, W, b, Wc, bc


def regularize_l1(W, G, l1, lr):
    """cu::RegularizeL1 CPU branch, src/aslp-cudamatrix/cu-math.cc:53-75."""
    W, G = np.array(W, F), np.array(G, F)
    l1s = np.where(W < 0, -F(l1), F(l1)).astype(F)
    after = W - F(lr) * G - l1s
    flip = ((after > 0) != (W > 0)) & (W != 0)
    keep = (~flip) & (W != 0)
    W[keep] = (W - l1s)[keep]
    W[flip] = 0
    G[flip] = 0
    return W, G


def max_norm_rows(W, max_norm):
    """nnet-affine-transform.h:232-243."""
    W = np.asarray(W, F)
    l2 = np.sqrt((W * W).sum(axis=1, dtype=F))
    scl = np.maximum(l2 * F(1.0 / max_norm), F(1))
    return (W * (F(1) / scl)[:, None]).astype(F)


# ----------------------------------------------------------------------------- loss
def xent(y, targets, frame_w):
    """Xent::Eval, src/aslp-nnet/nnet-loss.cc:63-129.  targets dense [rows, K].
    Returns diff and (cross_entropy, entropy, likelihood, correct, frames)."""
    y, t, fw = np.asarray(y, F), np.asarray(targets, F), np.asarray(frame_w, F)
    w = (fw * t.sum(axis=1, dtype=F)).astype(F)                    # :76-78
    diff = ((y - t) * w[:, None]).astype(F)                        # :83-85
    correct = float((w * (y.argmax(axis=1) == t.argmax(axis=1))).sum(dtype=np.float64))   # :88-91
    ce = -float((np.log(y + F(1e-20), dtype=F) * t * w[:, None]).sum(dtype=np.float64))   # :93-98
    en = -float((np.log(t + F(1e-20), dtype=F) * t * w[:, None]).sum(dtype=np.float64))   # :100-105
    lk = float((y * t * w[:, None]).sum(dtype=np.float64))                               # :107-110
    return diff, (ce, en, lk, correct, float(w.sum(dtype=np.float64)))


# ----------------------------------------------------------------------------- splice / rowconv
def splice_fwd(x, offsets):
    """cu::Splice CPU branch, src/aslp-cudamatrix/cu-math.cc:153-166."""
    x = np.asarray(x, F)
    n = x.shape[0]
    idx = np.clip(np.arange(n)[:, None] + np.asarray(offsets)[None, :], 0, n - 1)
    return x[idx].reshape(n, -1).astype(F)


def splice_bwd(dy, offsets, dim):
    """Splice::BackpropagateFnc, src/aslp-nnet/nnet-various.h:143-175 (gathers with the same
    clamp(t + offset) index as the forward pass -- a quirk that is kept)."""
    dy = np.asarray(dy, F)
    n = dy.shape[0]
    out = np.zeros((n, dim), F)
    for c, off in enumerate(offsets):
        idx = np.clip(np.arange(n) + off, 0, n - 1)
        out += dy[idx, c * dim:(c + 1) * dim]
    return out


def rowconv_fwd(x, w, S, lens):
    """RowConvolution::PropagateFnc, src/aslp-nnet/nnet-row-convolution.cc:90-118."""
    x, w = np.asarray(x, F), np.asarray(w, F)
    T, D, Fc = x.shape[0] // S, x.shape[1], w.shape[1] - 1
    out = np.zeros_like(x)
    for s in range(S):
        for t in range(lens[s]):
            acc = np.zeros(D, F)
            for k in range(Fc + 1):
                acc += w[:, k] * x[min(t + k, lens[s] - 1) * S + s]
            out[t * S + s] = acc
    return out


def rowconv_bwd(x, dy, w, S, lens):
    """RowConvolution::BackpropagateFnc, nnet-row-convolution.cc:120-159 -> (dx, w_diff)."""
    x, dy, w = np.asarray(x, F), np.asarray(dy, F), np.asarray(w, F)
    T, D, Fc = x.shape[0] // S, x.shape[1], w.shape[1] - 1
    dx = np.zeros_like(x)
    wd = np.zeros_like(w)
    for s in range(S):
        L = lens[s]
        buf = np.zeros((T + Fc, D), F)
        for t in range(L):
            for k in range(Fc + 1):
                buf[t + k] += w[:, k] * dy[t * S + s]
                wd[:, k] += x[min(t + k, L - 1) * S + s] * dy[t * S + s]
        for t in range(L):
            dx[t * S + s] = buf[t]
    return dx, wd


# ----------------------------------------------------------------------------- batch norm
def bn_fwd_train(x, scale, shift, var_floor=1e-7):
    """BatchNormalization::PropagateFnc, src/aslp-nnet/nnet-batch-normalization.h:176-220.
    Returns out, xhat, mean, inv_std, sum_x (f64), sum_x2 (f64)."""
    x = np.asarray(x, F)
    n = x.shape[0]
    mean = (x.sum(axis=0, dtype=F) * F(1.0 / n)).astype(F)
    xc = x - mean
    var = ((xc * xc).sum(axis=0, dtype=F) * F(1.0 / n)).astype(F)
    inv_std = (F(1) / np.sqrt(var + F(var_floor), dtype=F)).astype(F)
    xhat = (xc * inv_std).astype(F)
    out = (xhat * np.asarray(scale, F) + np.asarray(shift, F)).astype(F)
    return out, xhat, mean, inv_std, x.astype(np.float64).sum(axis=0), (x * x).astype(np.float64).sum(axis=0)


def bn_fwd_eval(x, scale, shift, mean, inv_std):
    """FeedforwardFnc global-stats branch, nnet-batch-normalization.h:167-174."""
    x = np.asarray(x, F)
    return (((x - mean) * inv_std) * np.asarray(scale, F) + np.asarray(shift, F)).astype(F)


def bn_bwd(x, xhat, dy, scale, mean, inv_std, mmt, dscale, dshift):
    """BackpropagateFnc, nnet-batch-normalization.h:222-277 -> (dx, dscale', dshift')."""
    x, xhat, dy, scale = (np.asarray(v, F) for v in (x, xhat, dy, scale))
    n = x.shape[0]
    dscale = (F(mmt) * np.asarray(dscale, F) + (xhat * dy).sum(axis=0, dtype=F)).astype(F)
    dshift = (F(mmt) * np.asarray(dshift, F) + dy.sum(axis=0, dtype=F)).astype(F)
    g = dy * scale
    dvar_c = (F(-0.5) * inv_std ** 3).astype(F)
    xc = x - mean
    dvar = (xc * g * dvar_c).sum(axis=0, dtype=F)
    dmean = (-(g * inv_std)).sum(axis=0, dtype=F)
    bufE = xc * F(2.0 / n) * dvar
    dmean = dmean - bufE.sum(axis=0, dtype=F)
    dx = g * inv_std + bufE + F(1.0 / n) * dmean
    return dx.astype(F), dscale, dshift


# ----------------------------------------------------------------------------- FSMN
def fsmn_fwd(x, coef, P, Fu):
    """CompactFsmn::PropagateFnc, src/aslp-nnet/nnet-cfsmn-component.h:169-198
    (AddConvMatMatElements + AddRowSumMat, cu-matrix.cc:3010-3072)."""
    x, coef = np.asarray(x, F), np.asarray(coef, F)
    T, D = x.shape
    C = P + Fu + 1
    pad = np.zeros((T + C - 1, D), F)
    pad[P:P + T] = x
    out = x.copy()
    for c in range(C):
        out += coef[c] * pad[c:c + T]
    return out


def fsmn_bwd(x, dy, coef, P, Fu, clip=0.0):
    """CompactFsmn::BackpropagateFnc, nnet-cfsmn-component.h:200-258 -> (dx, coef_corr)."""
    x, dy, coef = np.asarray(x, F), np.asarray(dy, F), np.asarray(coef, F)
    T, D = x.shape
    C = P + Fu + 1
    pad = np.zeros((T + C - 1, D), F)
    pad[P:P + T] = x
    corr = np.stack([(pad[c:c + T] * dy).sum(axis=0, dtype=F) for c in range(C)]).astype(F)   # beta = 0 (:224)
    padd = np.zeros((T + C - 1, D), F)
    padd[Fu:Fu + T] = dy
    rev = coef[::-1]
    dx = dy.copy()
    for c in range(C):
        dx += rev[c] * padd[c:c + T]
    if clip > 0:
        corr = np.clip(corr, -clip, clip)
    return dx, corr


# ----------------------------------------------------------------------------- LSTM family
def lstm_dir_fwd(gifo_x, w_r, w_rm, peep, state0, T, S, C, R, reverse=False, seq_len=None, clip=50.0):
    """One direction of the LSTM recurrence in the reference buffer layout.
    Lstm (nnet-recurrent-component.cc:235-335), LstmProjectedStreams (nnet-lstm-projected-streams.h:313-433),
    BLstmProjectedStreamsLC (nnet-blstm-projected-streams-lc.h:552-628 forward dir, :632-717 backward dir),
    BLstmProjectedStreams seq-length zeroing (nnet-blstm-projected-streams.h:654-657).
    gifo_x: [T*S, 4C] = x W_x^T + bias.  state0: [S, 7C+R] or None.
    Returns buf [(T+2)*S, 7C+R] with columns [g i f o c h m r]."""
    W = 7 * C + R
    buf = np.zeros(((T + 2) * S, W), F)
    buf[S:(T + 1) * S, :4 * C] = np.asarray(gifo_x, F)
    b0 = (T + 1) if reverse else 0
    if state0 is not None:
        buf[b0 * S:(b0 + 1) * S] = np.asarray(state0, F)
    w_r = np.asarray(w_r, F)
    pi, pf, po = (np.asarray(p, F) for p in peep)
    rc = slice(7 * C, 7 * C + R) if R > 0 else slice(6 * C, 7 * C)
    order = range(T, 0, -1) if reverse else range(1, T + 1)
    for t in order:
        tp = t + 1 if reverse else t - 1
        y = buf[t * S:(t + 1) * S]
        yp = buf[tp * S:(tp + 1) * S]
        gifo = y[:, :4 * C] + yp[:, rc] @ w_r.T
        cprev = yp[:, 4 * C:5 * C]
        g = tanh(gifo[:, :C])
        i = sigmoid(gifo[:, C:2 * C] + cprev * pi)
        f = sigmoid(gifo[:, 2 * C:3 * C] + cprev * pf)
        c = np.clip(g * i + cprev * f, -clip, clip).astype(F)
        h = tanh(c)
        o = sigmoid(gifo[:, 3 * C:4 * C] + c * po)
        m = (h * o).astype(F)
        y[:, :C], y[:, C:2 * C], y[:, 2 * C:3 * C], y[:, 3 * C:4 * C] = g, i, f, o
        y[:, 4 * C:5 * C], y[:, 5 * C:6 * C], y[:, 6 * C:7 * C] = c, h, m
        if R > 0:
            y[:, 7 * C:] = m @ np.asarray(w_rm, F).T
        if seq_len is not None:
            for s in range(S):
                if t > seq_len[s]:
                    y[s] = 0
    return buf


def lstm_dir_bwd(buf, out_diff, w_r, w_rm, peep, T, S, C, R, reverse=False):
    """BPTT of one direction (reference 'version 1' exact gradients):
    nnet-blstm-projected-streams-lc.h:763-835 (forward dir), :862-960 (backward dir);
    nnet-recurrent-component.cc:337-440.  out_diff: [T*S, R or C].
    Returns dbuf [(T+2)*S, 7C+R] with columns [dg di df do dc dh dm dr]."""
    W = 7 * C + R
    d = np.zeros(((T + 2) * S, W), F)
    oc = slice(7 * C, 7 * C + R) if R > 0 else slice(6 * C, 7 * C)
    d[S:(T + 1) * S, oc] = np.asarray(out_diff, F)
    w_r = np.asarray(w_r, F)
    pi, pf, po = (np.asarray(p, F) for p in peep)
    order = range(1, T + 1) if reverse else range(T, 0, -1)
    for t in order:
        tn = t - 1 if reverse else t + 1
        tp = t + 1 if reverse else t - 1
        y, yn, yp = buf[t * S:(t + 1) * S], buf[tn * S:(tn + 1) * S], buf[tp * S:(tp + 1) * S]
        dt, dn = d[t * S:(t + 1) * S], d[tn * S:(tn + 1) * S]
        yg, yi, yf, yo, yh = y[:, :C], y[:, C:2 * C], y[:, 2 * C:3 * C], y[:, 3 * C:4 * C], y[:, 5 * C:6 * C]
        if R > 0:
            dr = dt[:, 7 * C:] + dn[:, :4 * C] @ w_r
            dt[:, 7 * C:] = dr
            dm = dr @ np.asarray(w_rm, F)
        else:
            dm = dt[:, 6 * C:7 * C] + dn[:, :4 * C] @ w_r
        dh = (F(1) - yh * yh) * (dm * yo)
        do = yo * (F(1) - yo) * (dm * yh)
        dc = dh + dn[:, 4 * C:5 * C] * yn[:, 2 * C:3 * C] + dn[:, C:2 * C] * pi + dn[:, 2 * C:3 * C] * pf + do * po
        df = yf * (F(1) - yf) * (dc * yp[:, 4 * C:5 * C])
        di = yi * (F(1) - yi) * (dc * yg)
        dg = (F(1) - yg * yg) * (dc * yi)
        dt[:, :C], dt[:, C:2 * C], dt[:, 2 * C:3 * C], dt[:, 3 * C:4 * C] = dg, di, df, do
        dt[:, 4 * C:5 * C], dt[:, 5 * C:6 * C], dt[:, 6 * C:7 * C] = dc, dh, dm
    return d


def lstm_dir_wgrads(buf, dbuf, x, T, S, C, R, mmt, corr, reverse=False, clip=0.0):
    """Chunk weight gradients with momentum and elementwise clip (lc.h:981-1017 fwd dir, :1022-1057 bwd dir;
    nnet-recurrent-component.cc:444-477).  corr: dict with keys w_x, w_r, bias, pi, pf, po, (w_rm)."""
    x = np.asarray(x, F)
    rows = slice(S, (T + 1) * S)
    prev = slice(2 * S, (T + 2) * S) if reverse else slice(0, T * S)
    dg = dbuf[rows, :4 * C]
    rc = slice(7 * C, 7 * C + R) if R > 0 else slice(6 * C, 7 * C)
    out = {}
    out["w_x"] = F(mmt) * corr["w_x"] + dg.T @ x
    out["w_r"] = F(mmt) * corr["w_r"] + dg.T @ buf[prev, rc]
    out["bias"] = F(mmt) * corr["bias"] + dg.sum(axis=0, dtype=F)
    out["pi"] = F(mmt) * corr["pi"] + (dbuf[rows, C:2 * C] * buf[prev, 4 * C:5 * C]).sum(axis=0, dtype=F)
    out["pf"] = F(mmt) * corr["pf"] + (dbuf[rows, 2 * C:3 * C] * buf[prev, 4 * C:5 * C]).sum(axis=0, dtype=F)
    out["po"] = F(mmt) * corr["po"] + (dbuf[rows, 3 * C:4 * C] * buf[rows, 4 * C:5 * C]).sum(axis=0, dtype=F)
    if R > 0:
        out["w_rm"] = F(mmt) * corr["w_rm"] + dbuf[rows, 7 * C:].T @ buf[rows, 6 * C:7 * C]
    if clip > 0:
        out = {k: np.clip(v, -clip, clip) for k, v in out.items()}
    return {k: v.astype(F) for k, v in out.items()}


# ----------------------------------------------------------------------------- GRU
def gru_fwd(zrm_x, w_zr_h, w_m_g, state0, T, S, H):
    """GruStreams::PropagateFnc, src/aslp-nnet/nnet-gru-streams.h:238-320.  buf columns [z r m g h]."""
    buf = np.zeros(((T + 2) * S, 5 * H), F)
    buf[S:(T + 1) * S, :3 * H] = np.asarray(zrm_x, F)
    if state0 is not None:
        buf[:S] = np.asarray(state0, F)
    w_zr_h, w_m_g = np.asarray(w_zr_h, F), np.asarray(w_m_g, F)
    for t in range(1, T + 1):
        y, yp = buf[t * S:(t + 1) * S], buf[(t - 1) * S:t * S]
        hp = yp[:, 4 * H:]
        zr = sigmoid(y[:, :2 * H] + hp @ w_zr_h.T)
        z, r = zr[:, :H], zr[:, H:]
        g = (r * hp).astype(F)
        m = tanh(y[:, 2 * H:3 * H] + g @ w_m_g.T)
        h = (hp - hp * z + z * m).astype(F)
        y[:, :H], y[:, H:2 * H], y[:, 2 * H:3 * H], y[:, 3 * H:4 * H], y[:, 4 * H:] = z, r, m, g, h
    return buf


def gru_bwd(buf, out_diff, w_zr_h, w_m_g, T, S, H):
    """GruStreams::BackpropagateFnc, nnet-gru-streams.h:322-400.  dbuf columns [dz dr dm dg dh]."""
    d = np.zeros(((T + 2) * S, 5 * H), F)
    d[S:(T + 1) * S, 4 * H:] = np.asarray(out_diff, F)
    w_zr_h, w_m_g = np.asarray(w_zr_h, F), np.asarray(w_m_g, F)
    for t in range(T, 0, -1):
        y, yn, yp = buf[t * S:(t + 1) * S], buf[(t + 1) * S:(t + 2) * S], buf[(t - 1) * S:t * S]
        dt, dn = d[t * S:(t + 1) * S], d[(t + 1) * S:(t + 2) * S]
        dh = dt[:, 4 * H:] + dn[:, :2 * H] @ w_zr_h + dn[:, 4 * H:] - dn[:, 4 * H:] * yn[:, :H] + dn[:, 3 * H:4 * H] * yn[:, H:2 * H]
        z, r, m, hp = y[:, :H], y[:, H:2 * H], y[:, 2 * H:3 * H], yp[:, 4 * H:]
        dm = (F(1) - m * m) * (dh * z)
        dg = dm @ w_m_g
        dr = r * (F(1) - r) * (dg * hp)
        dz = z * (F(1) - z) * (dh * m - dh * hp)
        dt[:, :H], dt[:, H:2 * H], dt[:, 2 * H:3 * H], dt[:, 3 * H:4 * H], dt[:, 4 * H:] = dz, dr, dm, dg, dh
    return d


# ----------------------------------------------------------------------------- aslp-parallel
def bsp_sync(params_per_rank, frames_per_rank):
    """BspWorker::Synchronize, src/aslp-parallel/bsp-worker.cc:33-58 (frame-weighted MODEL average);
    MPI_Allreduce(SUM) restated as a sum over replicas."""
    tot = float(sum(frames_per_rank))
    if tot == 0:
        return None
    acc = sum(np.asarray(p, F) * F(f / tot) for p, f in zip(params_per_rank, frames_per_rank))
    return acc.astype(F)


def bmuf_sync(w_per_rank, w_prev, delta_prev, momentum, lr):
    """BmufWorker::Synchronize, src/aslp-parallel/bmuf-worker.cc:37-68 -> (w, w_prev', delta_prev')."""
    g = sum(np.asarray(w, F) - np.asarray(w_prev, F) for w in w_per_rank).astype(F)      # SUM, not mean
    delta = (F(momentum) * np.asarray(delta_prev, F) + F(1 - momentum) * F(lr) * g).astype(F)
    w = (np.asarray(w_prev, F) + delta).astype(F)
    return w, w.copy(), delta


def sod_optimize(opt, w, g, s1, s2, lr, p1, p2, step, floor=1e-8):
    """Optimizer family, src/aslp-parallel/optimizer.h:36-159, applied by SodWorker (sod-worker.cc:46-60)."""
    w, g, s1, s2 = (np.array(v, F) for v in (w, g, s1, s2))
    fl = F(floor)
    if opt == "sgd":
        w -= F(lr) * g
    elif opt == "momentum":
        s1 = F(p1) * s1 + F(lr) * g
        w -= s1
    elif opt == "adagrad":
        s1 = s1 + g * g
        w -= F(lr) * g / np.sqrt(np.maximum(s1, fl))
    elif opt == "rmsprop":
        s1 = F(0.9) * s1 + F(0.1) * g * g
        w -= F(lr) * g / np.sqrt(np.maximum(s1, fl))
    elif opt == "adadelta":
        s1 = F(p1) * s1 + F(1 - p1) * g * g
        upd = (F(1) / np.sqrt(np.maximum(s1, fl))) * np.sqrt(np.maximum(s2, fl)) * g
        w -= upd
        s2 = F(p1) * s2 + F(1 - p1) * upd * upd
    elif opt == "adam":
        s1 = F(p1) * s1 + F(1 - p1) * g
        s2 = F(p2) * s2 + F(1 - p2) * g * g
        c1, c2 = F(1.0 / (1 - p1 ** step)), F(1.0 / (1 - p2 ** step))
        w -= F(lr) * c1 * s1 / np.sqrt(np.maximum(s2 * c2, fl))
    else:
        raise ValueError(opt)
    return w.astype(F), s1.astype(F), s2.astype(F)


# ---------------------------------------------------------------- Eesen CTC on probabilities.  The reference implementation
# is GPU-only and has no tests; this follows the kernel sources line by line.  PINNED on the GPU box: tests/test_gpu_ctc_eesen.py
# runs the reference's own kernels (cu-kernels.cu compiled for sm_100a into oracle/_ref/libref_cukernels.so) and compares this
# restatement -- and aslp_ctc_eesen -- with them; it is also cross-checked against warp-ctc, whose CPU path is pinned by its own
# known-answer tests.
_LZ, _LINF, _EXPLIM, _FMAX = F(-1e30), F(1e30), F(88.722839), F(3.4028235e38)


# ---------------------------------------------------------------- async parameter-server modes (aslp-parallel)
# The reference has no tests for them and needs MPI; pinned, like the synchronous workers, against the reference's own server and
# worker classes run over oracle/stub/mpi.h (tests/test_cpu_oracle_pinning.py::test_async_restatements_match_the_reference_servers).
# One call = what the server and the worker do for ONE kMsgSynchronize, in the order the server receives the messages.

def easgd_exchange(w_worker, w_server, alpha):
    """easgd-worker.cc:58-62 and easgd-server.cc:78-83: both sides move towards the OTHER side's pre-update model."""
    a = F(alpha)
    new_worker = (a * w_server + (F(1) - a) * w_worker).astype(F)
    new_server = (a * w_worker + (F(1) - a) * w_server).astype(F)
    return new_worker, new_server


def asgd_update(w_worker, w_prev_worker, w_server, alpha):
    """asgd-worker.cc:40-44 + asgd-server.cc:88-92: delta = w - w_prev; server += alpha * delta; the worker restarts from
    the server's new model (outside the periodic barrier)."""
    delta = (w_worker - w_prev_worker).astype(F)
    new_server = (w_server + F(alpha) * delta).astype(F)
    return new_server.copy(), new_server


def masgd_update(w_worker, w_prev_worker, w_server, diff_k, momentum):
    """masgd-server.cc:121-123 (LMASGD): d_k = momentum * d_k + delta ; server += d_k."""
    delta = (w_worker - w_prev_worker).astype(F)
    d = (delta + F(momentum) * diff_k).astype(F)
    new_server = (w_server + d).astype(F)
    return new_server.copy(), new_server, d


def _add_ab(a, b):      # ctc-utils.h:60-65
    return _LZ if (a == _LZ or b == _LZ) else F(a + b)


def _sub_ab(a, b):      # ctc-utils.h:67-73
    if a == _LZ:
        return _LZ
    if b == _LZ:
        return _LINF
    return F(a - b)


def _exp_a(a):          # ctc-utils.h:41-47
    if a <= _LZ:
        return F(0)
    if a >= _EXPLIM:
        return _FMAX
    return F(np.exp(F(a)))


def _log_a_plus_b(a, b):   # ctc-utils.h:76-82
    if b < a:
        return _add_ab(a, F(np.log(F(1) + _exp_a(_sub_ab(b, a)))))
    return _add_ab(b, F(np.log(F(1) + _exp_a(_sub_ab(a, b)))))


def ctc_eesen(probs, labels, seq_len, T, S):
    """Ctc::EvalParallel (src/aslp-nnet/ctc-loss.cc:115-189) with the kernels _compute_ctc_alpha/beta/error_multiple_sequence
    (src/aslp-cudamatrix/cu-kernels.cu:3276-3315, 3391-3451, 3512-3534).  probs [T*S, K] stream-interleaved softmax outputs.
    Returns (pzx[S] = log p(z|x), diff [T*S, K] w.r.t. the pre-softmax activations, before the +-1 clip)."""
    probs = np.asarray(probs, F)
    K = probs.shape[1]
    Lmax = max(len(l) for l in labels)
    E = 2 * Lmax + 1
    lab = -np.ones((S, E), np.int64)
    for s, l in enumerate(labels):
        for i, c in enumerate(l):
            lab[s, 2 * i] = 0
            lab[s, 2 * i + 1] = c
        lab[s, 2 * len(l)] = 0
    with np.errstate(divide="ignore"):
        logp = np.log(probs).astype(F)
    alpha = np.full((T * S, E), _LZ, F)
    beta = np.full((T * S, E), _LZ, F)
    for t in range(T):
        for s in range(S):
            if t >= seq_len[s]:
                continue
            r, rp = t * S + s, (t - 1) * S + s
            for j in range(E):
                c = lab[s, j]
                if c == -1:
                    continue
                lp = logp[r, c]
                if t == 0:
                    alpha[r, j] = lp if j < 2 else _LZ
                elif j > 1:
                    tmp = _log_a_plus_b(alpha[rp, j - 1], alpha[rp, j])
                    if not (j % 2 == 0 or lab[s, j - 2] == c):
                        tmp = _log_a_plus_b(alpha[rp, j - 2], tmp)
                    alpha[r, j] = _add_ab(lp, tmp)
                elif j == 1:
                    alpha[r, j] = _add_ab(lp, _log_a_plus_b(alpha[rp, 0], alpha[rp, 1]))
                else:
                    alpha[r, j] = _add_ab(lp, alpha[rp, 0])
    for t in range(T - 1, -1, -1):
        for s in range(S):
            if t >= seq_len[s]:
                continue
            r, rn = t * S + s, (t + 1) * S + s
            ll = 2 * len(labels[s]) + 1
            for j in range(E):
                c = lab[s, j]
                if c == -1:
                    continue
                lp = logp[r, c]
                if t == seq_len[s] - 1:
                    beta[r, j] = lp if j > ll - 3 else _LZ
                elif j < ll - 2:
                    tmp = _log_a_plus_b(beta[rn, j + 1], beta[rn, j])
                    if not (j % 2 == 0 or lab[s, j + 2] == c):
                        tmp = _log_a_plus_b(beta[rn, j + 2], tmp)
                    beta[r, j] = _add_ab(lp, tmp)
                elif j == ll - 2:
                    beta[r, j] = _add_ab(lp, _log_a_plus_b(beta[rn, j + 1], beta[rn, j]))
                else:
                    beta[r, j] = _add_ab(lp, beta[rn, j])
    pzx = np.zeros(S, F)
    for s in range(S):
        ll = 2 * len(labels[s]) + 1
        r = (seq_len[s] - 1) * S + s
        t1, t2 = np.float64(alpha[r, ll - 1]), np.float64(alpha[r, ll - 2])      # LogAPlusB<double>: -1e30 is no sentinel there
        hi, lo = (t1, t2) if t1 > t2 else (t2, t1)
        pzx[s] = F(hi + np.log1p(np.exp(lo - hi)))
    err = np.zeros((T * S, K), F)
    for r in range(T * S):
        s, t = r % S, r // S
        if t >= seq_len[s]:
            continue
        for k in range(K):
            e = _LZ
            for j in range(E):
                if lab[s, j] == k:
                    e = _log_a_plus_b(e, _add_ab(alpha[r, j], beta[r, j]))
            y = probs[r, k]
            val = _exp_a(_sub_ab(e, _add_ab(pzx[s], _LZ if y == 0 else F(2) * F(np.log(y)))))
            err[r, k] = F(-1.0) * val
    err = (err * probs).astype(F)                                   # ctc_err_.MulElements(net_out)
    row_sum = err.sum(axis=1, dtype=F)
    diff = (err - probs * row_sum[:, None]).astype(F)
    return pzx, diff


# ---------------------------------------------------------------- ConvolutionalComponent / MaxPoolingComponent
def conv_column_map(in_dim, patch_dim, patch_step, patch_stride):
    """column_map_ of ConvolutionalComponent::PropagateFnc, src/aslp-nnet/nnet-convolutional-component.h:289-297."""
    num_splice = in_dim // patch_stride
    num_patches = 1 + (patch_stride - patch_dim) // patch_step
    return np.array([p * patch_step + s * patch_stride + d for p in range(num_patches) for s in range(num_splice) for d in range(patch_dim)],
                    np.int64), num_patches, num_splice * patch_dim


def conv_fwd(x, filters, bias, patch_dim, patch_step, patch_stride):
    """CopyCols + per-patch AddVecToRows / AddMatMat (:298-306).  Returns (out, vectorized_feature_patches)."""
    x = np.asarray(x, F)
    cmap, npatch, fd = conv_column_map(x.shape[1], patch_dim, patch_step, patch_stride)
    vec = x[:, cmap]
    nf = filters.shape[0]
    out = np.zeros((x.shape[0], npatch * nf), F)
    for p in range(npatch):
        out[:, p * nf:(p + 1) * nf] = bias[None, :] + vec[:, p * fd:(p + 1) * fd] @ filters.T
    return out.astype(F), vec


def conv_bwd(out_diff, filters, in_dim, patch_dim, patch_step, patch_stride):
    """per-patch AddMatMat into feature_patch_diffs_, then AddCols over the rearranged reverse map (:377-400): every input
    column adds its patch positions in ascending order of the forward index."""
    out_diff = np.asarray(out_diff, F)
    cmap, npatch, fd = conv_column_map(in_dim, patch_dim, patch_step, patch_stride)
    nf = filters.shape[0]
    pd = np.zeros((out_diff.shape[0], npatch * fd), F)
    for p in range(npatch):
        pd[:, p * fd:(p + 1) * fd] = out_diff[:, p * nf:(p + 1) * nf] @ filters
    in_diff = np.zeros((out_diff.shape[0], in_dim), F)
    for j, c in enumerate(cmap):                     # ascending j == the order of the AddCols passes for a given column
        in_diff[:, c] = in_diff[:, c] + pd[:, j]
    return in_diff, pd


def conv_grads(out_diff, vec, num_filters):
    """Update (:403-421): gradients summed over the patch positions."""
    npatch = out_diff.shape[1] // num_filters
    fd = vec.shape[1] // npatch
    fg = np.zeros((num_filters, fd), np.float64)
    bg = np.zeros(num_filters, np.float64)
    for p in range(npatch):
        dp = out_diff[:, p * num_filters:(p + 1) * num_filters].astype(np.float64)
        fg += dp.T @ vec[:, p * fd:(p + 1) * fd].astype(np.float64)
        bg += dp.sum(axis=0)
    return fg.astype(F), bg.astype(F)


def maxpool_fwd(x, pool_size, pool_step, pool_stride):
    """MaxPoolingComponent::PropagateFnc, src/aslp-nnet/nnet-max-pooling-component.h:100-114."""
    x = np.asarray(x, F)
    num_patches = x.shape[1] // pool_stride
    num_pools = 1 + (num_patches - pool_size) // pool_step
    out = np.full((x.shape[0], num_pools * pool_stride), F(-1e20), F)
    for q in range(num_pools):
        for r in range(pool_size):
            p = r + q * pool_step
            out[:, q * pool_stride:(q + 1) * pool_stride] = np.maximum(out[:, q * pool_stride:(q + 1) * pool_stride],
                                                                       x[:, p * pool_stride:(p + 1) * pool_stride])
    return out


def maxpool_bwd(x, out, out_diff, pool_size, pool_step, pool_stride):
    """BackpropagateFnc (:116-156): equality mask per (pool, member), summed, scaled by 1 / #pools containing the patch."""
    x, out, out_diff = np.asarray(x, F), np.asarray(out, F), np.asarray(out_diff, F)
    num_patches = x.shape[1] // pool_stride
    num_pools = 1 + (num_patches - pool_size) // pool_step
    in_diff = np.zeros_like(x)
    summands = [0] * num_patches
    for q in range(num_pools):
        for r in range(pool_size):
            p = r + q * pool_step
            sl_p, sl_q = slice(p * pool_stride, (p + 1) * pool_stride), slice(q * pool_stride, (q + 1) * pool_stride)
            in_diff[:, sl_p] = in_diff[:, sl_p] + out_diff[:, sl_q] * (x[:, sl_p] == out[:, sl_q]).astype(F)
            summands[p] += 1
    for p in range(num_patches):
        in_diff[:, p * pool_stride:(p + 1) * pool_stride] *= F(1.0 / summands[p])
    return in_diff
