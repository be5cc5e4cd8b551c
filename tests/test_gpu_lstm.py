"""GPU parity: persistent LSTM recurrence (aslp_lstm_seq_fwd / aslp_lstm_seq_bwd) vs the oracle's
restatement of the reference per-step loops (nnet-blstm-projected-streams-lc.h:552-960,
nnet-recurrent-component.cc:235-440, nnet-blstm-projected-streams.h:654-657)."""
import ctypes

import numpy as np
import pytest

from oracle import aslp_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star: per-frame outputs and gradients within 1e-4 relative (fp32 path)


def rel_err(got, want):
    return np.abs(got - want).max() / (np.abs(want).max() + 1e-30)


def make_dir(rng, T, S, C, R, D, scale=None):
    Rr = R if R > 0 else C
    # keep the random recurrent map contractive (spectral radius < 1): a chaotic net amplifies last-ulp
    # differences exponentially in T and says nothing about parity
    if scale is None:
        scale = min(0.3, 1.0 / np.sqrt(max(Rr, C)))
    p = {
        "w_x": (rng.uniform(-scale, scale, (4 * C, D))).astype(np.float32),
        "w_r": (rng.uniform(-scale, scale, (4 * C, Rr))).astype(np.float32),
        "bias": (rng.uniform(-scale, scale, 4 * C)).astype(np.float32),
        "pi": rng.uniform(-scale, scale, C).astype(np.float32),
        "pf": rng.uniform(-scale, scale, C).astype(np.float32),
        "po": rng.uniform(-scale, scale, C).astype(np.float32),
    }
    if R > 0:
        p["w_rm"] = rng.uniform(-scale, scale, (R, C)).astype(np.float32)
    return p


def run_case(T, S, C, R, ndirs, seed=0, with_state=True, seq_len=None, reverse_first=False, od_scale=None):
    import torch
    import kaldi_aslp_b200 as K
    from tests.gpu_utils import DMat, dvec, lib, ok, ptr, stream, sync, P
    rng = np.random.default_rng(seed)
    D = 12
    L = lib()
    W = 7 * C + R
    x = rng.standard_normal((T * S, D)).astype(np.float32)
    dirs_np, dev = [], []
    arr = (K.LstmDir * ndirs)()
    for di in range(ndirs):
        reverse = (di == 1) != reverse_first
        p = make_dir(rng, T, S, C, R, D)
        gifo_x = (x @ p["w_x"].T + p["bias"][None, :]).astype(np.float32)
        state0 = (rng.uniform(-0.5, 0.5, (S, W)).astype(np.float32) if (with_state and not reverse) else None)
        want = O.lstm_dir_fwd(gifo_x, p["w_r"], p.get("w_rm"), (p["pi"], p["pf"], p["po"]), state0, T, S, C, R,
                              reverse=reverse, seq_len=seq_len if reverse else None)
        # device buffer: gifo preactivations in rows [S,(T+1)S), boundary state, everything else zero
        buf0 = np.zeros(((T + 2) * S, W), np.float32)
        buf0[S:(T + 1) * S, :4 * C] = gifo_x
        if state0 is not None:
            buf0[:S] = state0
        dbuf = DMat(buf0)
        dw_r, dpi, dpf, dpo = DMat(p["w_r"]), dvec(p["pi"]), dvec(p["pf"]), dvec(p["po"])
        dw_rm = DMat(p["w_rm"]) if R > 0 else None
        dsl = dvec(seq_len, np.int32) if (seq_len is not None and reverse) else None
        a = arr[di]
        a.T, a.S, a.C, a.R, a.reverse = T, S, C, R, int(reverse)
        a.buf, a.ldb = dbuf.t.data_ptr(), dbuf.ld
        a.dbuf, a.lddb = None, 0
        a.w_r, a.ldwr = dw_r.t.data_ptr(), dw_r.ld
        a.w_rm, a.ldwrm = (dw_rm.t.data_ptr(), dw_rm.ld) if R > 0 else (None, 0)
        a.peep_i, a.peep_f, a.peep_o = dpi.data_ptr(), dpf.data_ptr(), dpo.data_ptr()
        a.seq_len_dev = dsl.data_ptr() if dsl is not None else None
        a.cell_clip = 50.0
        dirs_np.append((p, want, reverse, x))
        dev.append((dbuf, dw_r, dw_rm, dpi, dpf, dpo, dsl))
    wsb = L.aslp_lstm_workspace_bytes(T, S, C, R, ndirs, 0)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    ok(L.aslp_lstm_seq_fwd(stream(), ctypes.byref(arr), ndirs, ptr(ws), wsb))
    sync()
    errs = []
    for di in range(ndirs):
        got = dev[di][0].np()
        want = dirs_np[di][1]
        rows = slice(S, (T + 1) * S)
        errs.append(rel_err(got[rows], want[rows]))
    # ---- backward
    wsb2 = L.aslp_lstm_workspace_bytes(T, S, C, R, ndirs, 1)
    ws2 = torch.empty(wsb2 + 256, dtype=torch.uint8, device="cuda")
    keep = []
    berrs = []
    wants = []
    for di in range(ndirs):
        p, want, reverse, _ = dirs_np[di]
        od = rng.standard_normal((T * S, R if R > 0 else C)).astype(np.float32)
        if od_scale is not None:                         # per-(frame, cell) magnitudes spread over many decades
            od = (od * od_scale(rng, od.shape)).astype(np.float32)
        dwant = O.lstm_dir_bwd(want, od, p["w_r"], p.get("w_rm"), (p["pi"], p["pf"], p["po"]), T, S, C, R, reverse=reverse)
        d0 = np.zeros(((T + 2) * S, W), np.float32)
        oc = slice(7 * C, 7 * C + R) if R > 0 else slice(6 * C, 7 * C)
        d0[S:(T + 1) * S, oc] = od
        dd = DMat(d0)
        # feed the ORACLE forward activations so the backward check is independent of forward error
        fb = DMat(want)
        arr[di].buf, arr[di].ldb = fb.t.data_ptr(), fb.ld
        arr[di].dbuf, arr[di].lddb = dd.t.data_ptr(), dd.ld
        keep.append((dd, fb))
        wants.append(dwant)
    ok(L.aslp_lstm_seq_bwd(stream(), ctypes.byref(arr), ndirs, ptr(ws2), wsb2))
    sync()
    for di in range(ndirs):
        got = keep[di][0].np()
        rows = slice(S, (T + 1) * S)
        berrs.append(rel_err(got[rows], wants[di][rows]))
    return errs, berrs


@pytest.mark.parametrize("T,S,C,R,ndirs", [
    (5, 3, 8, 4, 1),          # tiny, projected, ragged S
    (7, 4, 8, 0, 1),          # no projection (Lstm)
    (20, 16, 64, 32, 2),      # bidirectional projected (LC-BLSTM shape class)
    (12, 100, 32, 0, 1),      # cfg2 stream count: several staging chunks
    (9, 20, 40, 24, 2),       # S not a multiple of 16
    (30, 16, 320, 320, 2),    # cfg3 layer geometry (C = R = 320, S = 16), short T
    (10, 8, 400, 16, 1),      # several cells per CTA (cb = 3), one projection row per CTA
    (10, 8, 64, 600, 1),      # several projection rows per CTA (rb = 5)
    (6, 16, 512, 0, 1),       # cfg2 cell count, no projection, cb = 4
    (30, 16, 320, 0, 2),      # cfg3 with the projection folded away: what the trainer launches (tensor-core form)
    (25, 40, 64, 0, 2),       # tensor-core form, several 16-stream chunks, ragged last chunk
    (8, 100, 512, 0, 1),      # cfg2 geometry: stream groups + chunks
    (9, 5, 24, 0, 2),         # tensor-core form, fewer streams than one chunk, C not a multiple of 16
])
@pytest.mark.parametrize("kernel", ["auto", "simt"])
def test_lstm_fwd_bwd_parity(T, S, C, R, ndirs, kernel, monkeypatch):
    # "auto" picks the mma.sync tensor-core recurrence whenever R == 0 and the shape fits, "simt" forces the FFMA form
    monkeypatch.setenv("ASLP_LSTM_KERNEL", kernel)
    errs, berrs = run_case(T, S, C, R, ndirs)
    assert max(errs) < RTOL, errs
    assert max(berrs) < RTOL, berrs


def test_blstm_seq_length_zeroing():
    # BLstmProjectedStreams: backward-direction rows with t > len[s] are zeroed (blstm-projected-streams.h:654-657)
    errs, berrs = run_case(10, 4, 16, 8, 2, seq_len=[10, 7, 3, 9], with_state=False)
    assert max(errs) < RTOL, errs
    assert max(berrs) < RTOL, berrs


def test_lstm_long_sequence_stays_in_tolerance():
    # error must not compound over many steps
    errs, berrs = run_case(300, 16, 64, 64, 2, seed=3)
    assert max(errs) < RTOL, errs
    assert max(berrs) < 5 * RTOL, berrs


@pytest.mark.parametrize("T,S,C,ndirs", [
    (300, 16, 64, 2),      # many times around the 4-slot partial-sum ring, 4 CTAs per chain
    (40, 12, 320, 2),      # cfg3 cell count, second stream group only half full (8 + 4 streams)
    (33, 40, 128, 1),      # five stream groups, one direction
    (7, 8, 384, 1),        # 24 CTAs per chain: three MMA tiles in every warp
    (6, 16, 512, 1),       # 32 CTAs per chain -> the wide form, 16 streams per chain (two n-tiles)
    (6, 72, 512, 1),       # wide form, 24 streams per chain (three n-tiles), three chains
    (8, 100, 512, 1),      # BASELINE cfg2 geometry: wide form, 32 streams per chain, last chain 4 streams
])
@pytest.mark.parametrize("form", ["transposed", "gather-all"])
def test_backward_recurrence_forms(T, S, C, ndirs, form, monkeypatch):
    """The two tensor-core backward kernels (lstm_bwd_t_kernel: partial d_m sums exchanged through a re-armed ring;
    lstm_bwd_mma_kernel: every CTA gathers all of dgifo) against the oracle on shapes the transposed form covers."""
    # "mma" makes a missing tensor-core plan an error; without the transposed form the cfg2-class shapes have none and
    # fall back to the SIMT kernel, so that arm runs with "auto"
    monkeypatch.setenv("ASLP_LSTM_KERNEL", "mma" if form == "transposed" else "auto")
    monkeypatch.setenv("ASLP_LSTM_BWD_T", "1" if form == "transposed" else "0")
    errs, berrs = run_case(T, S, C, 0, ndirs, seed=11)
    assert max(errs) < RTOL, errs
    assert max(berrs) < (5 * RTOL if T >= 300 else RTOL), berrs


@pytest.mark.parametrize("T,S,C", [(24, 16, 320), (12, 8, 64)])
def test_backward_tiny_and_widely_spread_derivatives(T, S, C, monkeypatch):
    """The transposed backward kernel scales every stream column of its fp16-split operand by a power of two taken from the
    column's largest |dgifo| over the CTA's 64 rows.  Late in a long utterance those derivatives are tiny (1e-9 is ordinary) and
    differ by many decades between neighbouring cells; the column maximum has to be exact in fp32 for that -- a maximum
    that went through fp16 on its way across lanes (flushes below 6e-8) picks a scale that overflows the real maximum.
    Output derivatives between 1e-13 and 1e-7 here: no Inf / NaN, and the usual bound relative to the largest value."""
    monkeypatch.setenv("ASLP_LSTM_KERNEL", "mma")
    monkeypatch.setenv("ASLP_LSTM_BWD_T", "1")
    spread = lambda rng, shape: 10.0 ** rng.uniform(-13.0, -7.0, size=shape)
    errs, berrs = run_case(T, S, C, 0, 2, seed=5, od_scale=spread)
    assert np.all(np.isfinite(berrs)) and max(berrs) < RTOL, berrs
