"""GPU parity: compute_ctc_loss (include/ctc.h, CTC_GPU) vs the C restatement of CpuCTC
(oracle/ctc_oracle.c) and warp-ctc's own known-answer cases (src/warp-ctc/tests/test_cpu.cpp)."""
import ctypes

import numpy as np
import pytest

from oracle import ctc_oracle as CO

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["fused", "warp", "block"])
def sweep_form(request, monkeypatch):
    """Every test runs with the three forms: the one-launch fused kernel (ctc_fused.cuh, the default up to 256 states and
    128 classes; larger problems fall through to the four-launch path), and the four-launch path with the warp-per-sweep
    and the one-state-per-thread CTA sweep kernels (ASLP_CTC_SWEEP=warp|block)."""
    monkeypatch.setenv("ASLP_CTC_SWEEP", request.param)
    return request.param


def grad_tol(costs, base=1e-4):
    """Gradient bound against the oracle: 1e-4, or 4 ulp(|cost|) where that is larger -- posteriors are
    exp(alpha + beta - log p - log Z), a difference of fp32 numbers of magnitude |cost|, so they carry a few ulp(|cost|) of
    absolute error in the reference itself.  The same bound for all three forms: the fused form rescales its rows to the
    reference's normaliser (the last alpha row) in its closing pass."""
    finite = np.abs(costs[np.isfinite(costs)])
    big = float(finite.max()) if finite.size else 1.0
    return max(base, 4 * float(np.spacing(np.float32(big))))


def gpu_ctc(acts, labels, ilen):
    import torch
    import kaldi_aslp_b200 as K
    from tests.gpu_utils import lib, ptr, stream, sync
    L = lib()
    acts = np.ascontiguousarray(acts, np.float32)
    maxT, mb, Kc = acts.shape
    flat = np.ascontiguousarray(np.concatenate([np.asarray(l, np.int32) for l in labels]), np.int32)
    llen = np.ascontiguousarray([len(l) for l in labels], np.int32)
    il = np.ascontiguousarray(ilen, np.int32)
    info = K.CtcComputeInfo(1, torch.cuda.current_stream().cuda_stream)
    size = ctypes.c_size_t(0)
    assert L.get_workspace_size(llen.ctypes.data, il.ctypes.data, Kc, mb, info, ctypes.addressof(size)) == 0
    ws = torch.empty(size.value + 256, dtype=torch.uint8, device="cuda")
    d_acts = torch.from_numpy(acts).cuda()
    d_grads = torch.zeros_like(d_acts)
    costs = np.zeros(mb, np.float32)
    rc = L.compute_ctc_loss(ptr(d_acts), ptr(d_grads), flat.ctypes.data, llen.ctypes.data, il.ctypes.data, Kc, mb,
                            costs.ctypes.data, ptr(ws), info)
    assert rc == 0, L.ctcGetStatusString(rc)
    sync()
    return costs, d_grads.cpu().numpy()


def gen_labels(rng, K, L, force_repeats=True):
    lab = list(rng.integers(1, K, size=L))
    if force_repeats and L >= 4:       # as tests/test.h:38-53 does: guarantee repeats
        lab[L // 2] = lab[L // 2 + 1]
        lab[L // 4] = lab[L // 4 + 1]
    return [int(v) for v in lab]


def test_ctc_small_known_answer():
    # src/warp-ctc/tests/test_cpu.cpp:12-67: T=2, K=5, labels {1,2}: score = p[t0][1] * p[t1][2]
    acts = np.array([0.1, 0.6, 0.1, 0.1, 0.1, 0.1, 0.1, 0.6, 0.1, 0.1], np.float32).reshape(2, 1, 5)
    e = np.exp(acts - acts.max(axis=2, keepdims=True))
    p = e / e.sum(axis=2, keepdims=True)
    expected = p[0, 0, 1] * p[1, 0, 2]
    costs, _ = gpu_ctc(acts, [[1, 2]], [2])
    assert abs(np.exp(-costs[0]) - expected) < 1e-6


def test_ctc_inf_activation():
    # test_cpu.cpp:69-122: a -1e30 activation gives an infinite cost and NaN-free gradients
    rng = np.random.default_rng(0)
    acts = rng.standard_normal((50, 1, 15)).astype(np.float32)
    acts[:, 0, 2] = -1e30
    labels = [gen_labels(rng, 15, 10)]
    labels[0][0] = 2
    costs, grads = gpu_ctc(acts, labels, [50])
    assert np.isinf(costs[0]) and costs[0] > 0
    assert not np.isnan(grads).any()
    c2, g2 = CO.cost_and_grad(acts, labels, [50])
    assert np.isinf(c2[0])
    assert np.abs(grads - g2).max() < 1e-5


@pytest.mark.parametrize("K,T,L,mb", [(20, 50, 15, 1), (5, 10, 5, 65), (72, 120, 30, 16), (72, 64, 31, 3)])
def test_ctc_vs_oracle(K, T, L, mb):
    rng = np.random.default_rng(K * 1000 + T)
    acts = rng.standard_normal((T, mb, K)).astype(np.float32)
    labels = [gen_labels(rng, K, L) for _ in range(mb)]
    ilen = [T] * mb
    costs, grads = gpu_ctc(acts, labels, ilen)
    c2, g2 = CO.cost_and_grad(acts, labels, ilen)
    assert np.allclose(costs, c2, rtol=1e-4, atol=1e-5), (costs, c2)       # north_star: CTC loss within 1e-4 relative
    assert np.abs(grads - g2).max() < grad_tol(c2)


def test_ctc_ragged_lengths_and_skips():
    # different input lengths, an utterance that is too short (L + repeats > T -> cost 0, zero grads), empty-ish labels
    rng = np.random.default_rng(7)
    K, maxT, mb = 24, 40, 6
    acts = rng.standard_normal((maxT, mb, K)).astype(np.float32)
    labels = [gen_labels(rng, K, 8), [3, 3, 3, 3, 3], gen_labels(rng, K, 12), [5], gen_labels(rng, K, 20), [7, 8]]
    ilen = [40, 6, 25, 1, 21, 40]     # utt 1: 5 labels + 4 repeats = 9 > 6 -> skipped ; utt 4: 20 labels + repeats > 21
    costs, grads = gpu_ctc(acts, labels, ilen)
    c2, g2 = CO.cost_and_grad(acts, labels, ilen)
    assert np.allclose(costs, c2, rtol=1e-4, atol=1e-5), (costs, c2)
    assert np.abs(grads - g2).max() < grad_tol(c2)
    assert costs[1] == 0.0 and np.all(grads[:, 1, :] == 0.0)
    for n in range(mb):
        assert np.all(grads[ilen[n]:, n, :] == 0.0)       # rows past the utterance end stay zero


def test_ctc_cfg3_geometry_many_utts_warp_path():
    # BASELINE cfg3 geometry (K=72, L=100 -> 201 states) with a minibatch large enough for the
    # warp-per-utterance path; T shortened so the oracle finishes in seconds
    rng = np.random.default_rng(11)
    K, T, L, mb = 72, 260, 100, 1300
    acts = rng.standard_normal((T, mb, K)).astype(np.float32)
    labels = [gen_labels(rng, K, L) for _ in range(mb)]
    ilen = [T] * mb
    costs, grads = gpu_ctc(acts, labels, ilen)
    sel = list(range(0, mb, 97))
    c2, g2 = CO.cost_and_grad(acts[:, sel, :], [labels[i] for i in sel], [T] * len(sel))
    assert np.allclose(costs[sel], c2, rtol=1e-4, atol=1e-5)
    # fp32 log-space: alpha+beta-logZ is a difference of numbers of magnitude |cost| ~ 1e3 whose ulp is 6e-5, so
    # posteriors (and the reference's own) carry ~4 ulp(|cost|) of absolute error; 1e-4 holds for short T (above)
    tol = grad_tol(c2)
    assert np.abs(grads[:, sel, :] - g2).max() < tol
    # size-independent property: every gradient row sums to ~0 (softmax - posterior, both sum to 1)
    # (up to the same fp32 log-space resolution: a few ulp(|cost|))
    assert np.abs(grads.sum(axis=2)).max() < 16 * float(np.spacing(np.float32(costs.max())))


def test_ctc_cpu_location_is_refused():
    import kaldi_aslp_b200 as K
    from tests.gpu_utils import lib
    L = lib()
    info = K.CtcComputeInfo(0, None)
    a = np.zeros((2, 1, 5), np.float32)
    g = np.zeros_like(a)
    lab = np.array([1], np.int32)
    ll = np.array([1], np.int32)
    il = np.array([2], np.int32)
    c = np.zeros(1, np.float32)
    ws = np.zeros(1024, np.float32)
    rc = L.compute_ctc_loss(a.ctypes.data, g.ctypes.data, lab.ctypes.data, ll.ctypes.data, il.ctypes.data, 5, 1,
                            c.ctypes.data, ws.ctypes.data, info)
    assert rc == 3      # CTC_STATUS_EXECUTION_FAILED: there is no CPU path


def test_ctc_more_than_256_states_uses_block_form():
    # 140 labels -> 281 states: past the warp form's 8 states per lane; must still match the oracle
    rng = np.random.default_rng(5)
    K, T, L, mb = 40, 330, 140, 3
    acts = rng.standard_normal((T, mb, K)).astype(np.float32)
    labels = [gen_labels(rng, K, L) for _ in range(mb)]
    costs, grads = gpu_ctc(acts, labels, [T] * mb)
    c2, g2 = CO.cost_and_grad(acts, labels, [T] * mb)
    assert np.allclose(costs, c2, rtol=1e-4, atol=1e-5)
    assert np.abs(grads - g2).max() < grad_tol(c2)


@pytest.mark.parametrize("L", [1, 2, 16, 17, 33, 64, 100, 127])
def test_ctc_state_counts_around_lane_boundaries(L):
    # 2L+1 states against the lanes x states-per-lane partition of the warp form (NS = 1 .. 8), with heavy label repeats
    rng = np.random.default_rng(100 + L)
    K, mb = 6, 4
    T = 2 * L + 9
    acts = rng.standard_normal((T, mb, K)).astype(np.float32)
    labels = [[int(v) for v in rng.integers(1, 3, size=L)] for _ in range(mb)]      # two symbols: many repeats
    ilen = [T, T - 3, T, max(1, T - 7)]
    costs, grads = gpu_ctc(acts, labels, ilen)
    c2, g2 = CO.cost_and_grad(acts, labels, ilen)
    assert np.allclose(costs, c2, rtol=1e-4, atol=1e-5), (costs, c2)
    assert np.abs(grads - g2).max() < grad_tol(c2)
