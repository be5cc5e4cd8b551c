"""Eesen-style CTC kernel (aslp_ctc_eesen) against (1) the reference's OWN CUDA kernels -- cu-kernels.cu compiled for sm_100a
from where it lies into oracle/_ref/libref_cukernels.so (oracle/Makefile) and driven here the way Ctc::EvalParallel drives them
(the reference has no CPU implementation and no tests of this path) --, (2) the numpy restatement
(oracle.aslp_oracle.ctc_eesen) and (3), where both apply, warp-ctc, whose CPU path is pinned by its own known-answer tests:
for p = softmax(x) the two formulations give the same cost and the same gradient w.r.t. x."""
import ctypes
import os

import numpy as np
import pytest
import torch

import kaldi_aslp_b200 as K
from oracle import aslp_oracle as O
from oracle import ctc_oracle

pytestmark = pytest.mark.gpu
P = ctypes.c_void_p


def run_kernel(probs, labels, seq_len, T, S):
    L = K.cuda_lib()
    Kc = probs.shape[1]
    Lexp = 2 * max(len(l) for l in labels) + 1
    lab = -np.ones((S, Lexp), np.int32)
    for s, l in enumerate(labels):
        for i, c in enumerate(l):
            lab[s, 2 * i] = 0
            lab[s, 2 * i + 1] = c
        lab[s, 2 * len(l)] = 0
    p_d = torch.from_numpy(np.ascontiguousarray(probs, np.float32)).cuda()
    d_d = torch.full_like(p_d, 7.0)                       # poison: every element must be written
    lab_d = torch.from_numpy(lab).cuda()
    len_d = torch.tensor(seq_len, dtype=torch.int32).cuda()
    pzx_d = torch.zeros(S, device="cuda")
    wsb = L.aslp_ctc_eesen_workspace_bytes(T, S, Kc, Lexp)
    ws = torch.empty(wsb + 16, dtype=torch.uint8, device="cuda")
    K.check(L.aslp_ctc_eesen(P(torch.cuda.current_stream().cuda_stream), P(d_d.data_ptr()), Kc, P(p_d.data_ptr()), Kc, T, S, Kc,
                             P(lab_d.data_ptr()), Lexp, P(len_d.data_ptr()), P(pzx_d.data_ptr()), P(ws.data_ptr()), wsb))
    torch.cuda.synchronize()
    return pzx_d.cpu().numpy(), d_d.cpu().numpy()


REF_CU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_cukernels.so")


class Dim3(ctypes.Structure):
    _fields_ = [("x", ctypes.c_uint), ("y", ctypes.c_uint), ("z", ctypes.c_uint)]


class MatrixDim(ctypes.Structure):                       # aslp-cudamatrix/cu-matrixdim.h:52-56
    _fields_ = [("rows", ctypes.c_int), ("cols", ctypes.c_int), ("stride", ctypes.c_int)]


def run_reference_kernels(probs, labels, seq_len, T, S):
    """Ctc::EvalParallel (src/aslp-nnet/ctc-loss.cc:115-185) with the reference's kernels doing what they do there:
    label expansion (:133-150), log of the net output (:153-154), alpha / beta set to log_zero then one
    ComputeCtcAlphaMSeq / ComputeCtcBetaMSeq launch per time step (:157-166; grid and block as cu-matrix.cc:2860-2862, 2927-2929),
    pzx from the last two alpha cells in double (:167-175), ComputeCtcErrorMSeq (:178-179; cu-matrix.cc:2996-2998), and the
    back-propagation through the softmax (:182-189).  Returns (pzx, diff) before the loss guards and the +-1 clip."""
    R = ctypes.CDLL(REF_CU)
    Kc = probs.shape[1]
    Lexp = 2 * max(len(l) for l in labels) + 1
    lab = -np.ones((S, Lexp), np.int32)
    for s, l in enumerate(labels):
        for i, c in enumerate(l):
            lab[s, 2 * i] = 0
            lab[s, 2 * i + 1] = c
        lab[s, 2 * len(l)] = 0
    lab_len = np.array([2 * len(l) + 1 for l in labels], np.int32)
    net_out = torch.from_numpy(np.ascontiguousarray(probs, np.float32)).cuda()
    logp = torch.log(net_out)                                      # CuMatrix::ApplyLog
    alpha = torch.full((T * S, Lexp), -1e30, device="cuda")
    beta = torch.full((T * S, Lexp), -1e30, device="cuda")
    lab_d = torch.from_numpy(lab.reshape(-1)).cuda()
    len_d = torch.tensor(seq_len, dtype=torch.int32).cuda()
    lablen_d = torch.from_numpy(lab_len).cuda()
    nb = lambda n: (n + 15) // 16                                   # n_blocks(n, CU2DBLOCK)
    block = Dim3(16, 16, 1)
    grid = Dim3(nb(S), nb(Lexp), 1)
    d_ab, d_p = MatrixDim(T * S, Lexp, Lexp), MatrixDim(T * S, Kc, Kc)
    torch.cuda.synchronize()                                       # the reference kernels run on the legacy default stream
    for t in range(T):
        R.cudaF_compute_ctc_alpha_multiple_sequence(grid, block, P(alpha.data_ptr()), S, t, d_ab, P(logp.data_ptr()), d_p, P(lab_d.data_ptr()), Lexp,
                                                    P(len_d.data_ptr()))
    for t in range(T - 1, -1, -1):
        R.cudaF_compute_ctc_beta_multiple_sequence(grid, block, P(beta.data_ptr()), S, t, d_ab, P(logp.data_ptr()), d_p, P(lab_d.data_ptr()), Lexp,
                                                   P(len_d.data_ptr()), P(lablen_d.data_ptr()))
    torch.cuda.synchronize()
    a = alpha.cpu().numpy().astype(np.float64)
    pzx = np.zeros(S, np.float32)
    for s in range(S):
        t1, t2 = a[(seq_len[s] - 1) * S + s, lab_len[s] - 1], a[(seq_len[s] - 1) * S + s, lab_len[s] - 2]
        hi, lo = max(t1, t2), min(t1, t2)
        pzx[s] = np.float32(hi + np.log(1.0 + np.exp(lo - hi)))    # (float)LogAPlusB((double)tmp1, (double)tmp2), ctc-utils.h:88-95
    pzx_d = torch.from_numpy(pzx).cuda()
    err = torch.zeros((T * S, Kc), device="cuda")
    R.cudaF_compute_ctc_error_multiple_sequence(Dim3(nb(T * S), nb(Kc), 1), block, P(err.data_ptr()), S, d_p, P(alpha.data_ptr()), P(beta.data_ptr()), d_ab,
                                                P(net_out.data_ptr()), P(lab_d.data_ptr()), Lexp, P(len_d.data_ptr()), P(pzx_d.data_ptr()))
    torch.cuda.synchronize()
    err = err * net_out                                            # MulElements
    row_sum = err.sum(dim=1, keepdim=True)                         # AddColSumMat
    diff = err - net_out * row_sum                                 # CopyFromMat; AddMat(-1, net_out .* row_sum)
    return pzx, diff.cpu().numpy()


def softmax(x):
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


@pytest.mark.parametrize("T,S,Kc,lens,lab_lens", [
    (12, 3, 8, [12, 9, 5], [3, 2, 4]),        # ragged frames and labels (K multiple of 4: rows are 16-byte aligned)
    (6, 1, 4, [6], [1]),
    (20, 4, 12, [20, 20, 17, 3], [5, 1, 6, 1]),
])
def test_eesen_ctc_matches_restated_reference(T, S, Kc, lens, lab_lens):
    rng = np.random.default_rng(T * 31 + S)
    probs = softmax(rng.standard_normal((T * S, Kc)).astype(np.float32) * 2)
    labels = [rng.integers(1, Kc, size=n).tolist() for n in lab_lens]
    labels[0][-1] = labels[0][-2] if len(labels[0]) > 1 else labels[0][-1]     # a repeated label: the blank between them is mandatory
    want_pzx, want_diff = O.ctc_eesen(probs, labels, lens, T, S)
    got_pzx, got_diff = run_kernel(probs, labels, lens, T, S)
    np.testing.assert_allclose(got_pzx, want_pzx, rtol=1e-5, atol=1e-5)
    assert np.max(np.abs(got_diff - want_diff)) <= 1e-4 * max(1e-3, np.max(np.abs(want_diff)))
    # rows past the end of a stream carry no error
    for s in range(S):
        for t in range(lens[s], T):
            assert not got_diff[t * S + s].any()


@pytest.mark.parametrize("T,S,Kc,lens,lab_lens", [
    (12, 3, 8, [12, 9, 5], [3, 2, 4]),
    (6, 1, 4, [6], [1]),
    (20, 4, 12, [20, 20, 17, 3], [5, 1, 6, 1]),
    (150, 8, 72, [150, 150, 141, 120, 97, 60, 33, 20], [20, 14, 25, 9, 11, 3, 6, 2]),      # the recipes' class count, ragged
])
def test_eesen_ctc_matches_the_reference_cuda_kernels(T, S, Kc, lens, lab_lens):
    """The reference's own kernels as the oracle (they are all the reference has for this path).  Both sides take the same fp32
    probabilities; pzx to 1e-5 relative (a few hundred fp32 log-adds each way), the derivative to 1e-4 of its largest element."""
    if not os.path.exists(REF_CU):
        pytest.skip("oracle/_ref/libref_cukernels.so is not built (make -C oracle, where /root/reference exists)")
    rng = np.random.default_rng(T * 17 + S)
    probs = softmax(rng.standard_normal((T * S, Kc)).astype(np.float32) * 2)
    labels = [rng.integers(1, Kc, size=n).tolist() for n in lab_lens]
    if len(labels[0]) > 1:
        labels[0][-1] = labels[0][-2]                      # a repeated label: the blank between them is mandatory
    want_pzx, want_diff = run_reference_kernels(probs, labels, lens, T, S)
    got_pzx, got_diff = run_kernel(probs, labels, lens, T, S)
    np.testing.assert_allclose(got_pzx, want_pzx, rtol=1e-5, atol=1e-5)
    assert np.max(np.abs(got_diff - want_diff)) <= 1e-4 * max(1e-3, np.max(np.abs(want_diff)))
    # ... and the numpy restatement agrees with the reference's kernels too (this is what pins it)
    o_pzx, o_diff = O.ctc_eesen(probs, labels, lens, T, S)
    np.testing.assert_allclose(o_pzx, want_pzx, rtol=1e-5, atol=1e-5)
    assert np.max(np.abs(o_diff - want_diff)) <= 1e-4 * max(1e-3, np.max(np.abs(want_diff)))


def test_eesen_and_warp_ctc_agree_on_cost_and_gradient():
    T, S, Kc, L = 25, 4, 8, 5
    rng = np.random.default_rng(5)
    x = rng.standard_normal((T * S, Kc)).astype(np.float32)
    labels = [rng.integers(1, Kc, size=L).tolist() for _ in range(S)]
    lens = [T] * S
    pzx, diff = run_kernel(softmax(x), labels, lens, T, S)
    costs, grads = ctc_oracle.cost_and_grad(x.reshape(T, S, Kc), labels, lens)        # warp-ctc semantics (softmax inside)
    np.testing.assert_allclose(-pzx, costs, rtol=1e-4)
    assert np.max(np.abs(diff - grads.reshape(T * S, Kc))) <= 2e-4 * np.max(np.abs(grads))


def test_eesen_ctc_host_class_through_the_net():
    """kaldi::aslp_nnet::Ctc behind the trainer loop body: for the golden BLSTM-CTC net, the loss derivative that reaches the
    Softmax component equals the restated Eesen error for the net's own outputs (clipped to +-1), the per-sequence objective
    equals -log p(z|x), and the Softmax back-propagation is the reference's pass-through copy."""
    import os
    from kaldi_aslp_b200 import nnet as NN
    from oracle import kaldi_io
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lc_blstm_ctc")
    x = kaldi_io.read(os.path.join(d, "input.mat"))
    labels = [list(map(int, l.split())) for l in open(os.path.join(d, "labels.txt"))]
    S, T = 3, x.shape[0] // 3
    net = NN.Nnet.read(os.path.join(d, "model.bin"))
    net.set_train_options(0.0, 0.0, 0.0, 0.0)                 # learn rate 0: parameters untouched
    ctc = NN.EesenCtc()
    obj = NN.train_step_ctc_eesen(net, ctc, x, [T] * S, labels, with_error_rate=True)
    ncomp = net.num_components
    probs = net.component_output(ncomp - 1, T * S, 8)
    got = net.component_out_diff(ncomp - 1, T * S, 8)
    want_pzx, want_diff = O.ctc_eesen(probs, labels, [T] * S, T, S)
    np.testing.assert_allclose(obj, -want_pzx, rtol=1e-5)
    want_diff = np.clip(want_diff, -1.0, 1.0)
    assert np.max(np.abs(got - want_diff)) <= 1e-4 * np.max(np.abs(want_diff))
    rep = ctc.report()
    assert "Obj(log[Pzx]) = " in rep and "TOKEN_ACCURACY" in rep
    assert abs(float(rep.split("Obj(log[Pzx]) = ")[1].split()[0]) - float(np.mean(obj))) < 1e-3
    net.close()


def test_eesen_ctc_trainer_runs_and_learns(tmp_path):
    """aslp-nnet-train-ctc-streams (no reference CPU build exists for it): two passes over the trainer-level CTC fixture; the
    objective of the second pass must be lower than the first, the bookkeeping must match the warp-ctc trainer's."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "kaldi-aslp_b200", "build", "bin", "aslp-nnet-train-ctc-streams")
    d = os.path.join(root, "tests", "golden", "cli_ctc")
    objs = []
    model = os.path.join(d, "init.nnet")
    for it in range(2):
        out = str(tmp_path / ("iter%d.nnet" % it))
        r = subprocess.run([exe, "--num-stream=3", "--learn-rate=0.05", "--momentum=0.9", "ark:" + os.path.join(d, "feats.ark"),
                            "ark:" + os.path.join(d, "labels.ark"), model, out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:]
        assert re.findall(r"Done (\d+) files, (\d+) with no targets", r.stdout)[-1] == ("7", "0")
        objs.append(float(re.findall(r"Obj\(log\[Pzx\]\) = ([\d.eE+-]+)", r.stdout)[-1]))
        model = out
    assert objs[1] < objs[0], objs
