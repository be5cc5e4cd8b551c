"""Eesen-style CTC kernel (aslp_ctc_eesen) against the restated reference (oracle.aslp_oracle.ctc_eesen, PARITY UNPINNED in the
reference itself: GPU-only code without tests) and, where both apply, against warp-ctc -- whose CPU path IS pinned by its own
known-answer tests: for p = softmax(x) the two formulations give the same cost and the same gradient w.r.t. x."""
import ctypes

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
