"""Parity at BASELINE.json's FULL sizes.

* test_cfg3_full_minibatch_matches_reference_cpu: one whole LC-BLSTM-CTC training minibatch of BASELINE config 3
  (3 x BLstmProjectedStreamsLC 320/dir + Affine 640->72 + Softmax, 16 utterances x 1000 frames, 100 labels each) on the
  GPU against the UNMODIFIED reference running on the box's CPU (oracle/_ref/ref_driver, a few seconds per step):
  outputs, CTC costs and every parameter after the update.
* size-independent properties at the same sizes (no reference needed): permuting the streams of a minibatch permutes
  the LSTM outputs bit for bit (streams are independent chains); the two directions are mirror images of each other;
  the CTC gradient rows sum to zero; GEMM column checksums against float64."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import kaldi_aslp_b200 as K
from oracle import kaldi_io

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
P = ctypes.c_void_p
T, S, D, KC, L = 1000, 16, 40, 72, 100


def cfg3_proto():
    lines = ["<NnetProto>"]
    for din in (D, 640, 640):
        lines.append("<BLstmProjectedStreamsLC> <InputDim> %d <OutputDim> 640 <CellDim> 320 <ParamScale> 0.01 <ClipGradient> 5" % din)
    lines.append("<AffineTransform> <InputDim> 640 <OutputDim> %d <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04" % KC)
    lines.append("<Softmax> <InputDim> %d <OutputDim> %d" % (KC, KC))
    lines.append("</NnetProto>")
    return "\n".join(lines) + "\n"


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="oracle/_ref/ref_driver not built")
def test_cfg3_full_minibatch_matches_reference_cpu(tmp_path):
    from kaldi_aslp_b200 import nnet as NN
    d = str(tmp_path)
    open(os.path.join(d, "proto.txt"), "w").write(cfg3_proto())
    env = dict(os.environ, OPENBLAS_NUM_THREADS=str(os.cpu_count() or 8))
    subprocess.check_call([DRIVER, "init", "proto.txt", "model.bin", "777", "1"], cwd=d, env=env, stderr=subprocess.DEVNULL)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((T * S, D)).astype(np.float32)
    labels = [rng.integers(1, KC, size=L).tolist() for _ in range(S)]
    kaldi_io.write_mat(os.path.join(d, "input.mat"), x)
    open(os.path.join(d, "labels.txt"), "w").write("\n".join(" ".join(map(str, l)) for l in labels) + "\n")
    spec = dict(momentum=0.9, iters=1, seq_lengths=",".join([str(T)] * S), norm_learn_rate=1e-3, input="input.mat", loss="ctc",
                labels="labels.txt")
    open(os.path.join(d, "spec.txt"), "w").write("".join("%s %s\n" % kv for kv in spec.items()))
    subprocess.check_call([DRIVER, "step", "model.bin", "spec.txt", "ref"], cwd=d, env=env, stderr=subprocess.DEVNULL, timeout=900)

    ref = lambda name: kaldi_io.read(os.path.join(d, "ref", name + ".iter0"))
    rel = lambda g, w: float(np.abs(g - w).max() / (np.abs(w).max() + 1e-30))

    # (1) the whole training minibatch through the trainer loop body (Propagate, WarpCtc::Eval, Backpropagate + updates)
    net = NN.Nnet.read(os.path.join(d, "model.bin"))
    net.set_train_options(0.0, 0.9, 0.0, 0.0)
    ctc = NN.WarpCtc()
    costs = np.asarray(NN.train_step_ctc(net, ctc, x, [T] * S, labels, norm_learn_rate=1e-3), np.float64)
    ncomp = net.num_components
    assert rel(net.component_output(ncomp - 1, T * S, KC), ref("out")) < 1e-4               # forward: 1e-4 (north_star)
    # CTC cost: 1e-4 relative (north_star); the reference's report carries the mean -log p(z|x) of the minibatch
    rep = open(os.path.join(d, "ref", "report.txt")).read()
    import re
    ref_obj = float(re.findall(r"Obj\(log\[Pzx\]\) = ([\d.eE+-]+)", rep)[-1])
    assert abs(costs.mean() - ref_obj) <= 1e-4 * ref_obj, (costs.mean(), ref_obj)
    # CTC gradient: grad = p - exp(alpha + beta - log p - logZ) with |logZ| ~ 3.7e3, i.e. an exponent known to ulp(3.7e3) =
    # 2.4e-4 in fp32 in BOTH implementations; they sum the log-domain terms in different orders, so the posteriors agree to a
    # few of those ulps, not to 1e-4.  Bound: 8 ulp of the cost (the kernel-level tests use the same yardstick).
    ulp = float(np.spacing(np.float32(costs.max())))
    ld = net.component_out_diff(ncomp - 1, T * S, KC)
    assert np.abs(ld - ref("loss_diff")).max() < 8 * ulp, (np.abs(ld - ref("loss_diff")).max(), ulp)
    net.close()

    # (2) the backward pass and the updates in isolation from that fp32 limit: feed the REFERENCE's loss derivative
    net = NN.Nnet.read(os.path.join(d, "model.bin"))
    net.set_seq_lengths([T] * S)
    net.set_train_options(1e-3 / (T * S), 0.9, 0.0, 0.0)                                    # learn_rate = norm_lr / valid frames
    out = net.propagate(x)
    assert rel(out, ref("out")) < 1e-4
    in_diff = net.backpropagate(ref("loss_diff"))
    assert rel(in_diff, ref("in_diff")) < 2e-4
    got_p, want_p = net.get_params(), ref("params")
    assert got_p.shape == want_p.shape == (6510792,)
    assert rel(got_p, want_p) < 1e-4
    # the update itself (what the step changed), not just the parameters it was added to
    init_p = NN.Nnet.read(os.path.join(d, "model.bin"))
    p0 = init_p.get_params()
    init_p.close()
    assert rel(got_p - p0, want_p - p0) < 2e-3
    net.close()


def lstm_run(x_gifo, w, peep, T_, S_, C, ndirs):
    """aslp_lstm_seq_fwd on prepared buffers (R = 0 form); returns the [g i f o c h m] buffers per direction"""
    Lb = K.cuda_lib()
    W = 7 * C
    arr = (K.LstmDir * ndirs)()
    keep, bufs = [], []
    for dd in range(ndirs):
        buf = torch.zeros(((T_ + 2) * S_, W), device="cuda")
        buf[S_:(T_ + 1) * S_, :4 * C] = torch.from_numpy(x_gifo[dd]).cuda()
        wr = torch.from_numpy(w[dd]).cuda()
        pe = [torch.from_numpy(p).cuda() for p in peep[dd]]
        a = arr[dd]
        a.T, a.S, a.C, a.R, a.reverse = T_, S_, C, 0, dd
        a.buf, a.ldb, a.dbuf, a.lddb = buf.data_ptr(), W, None, 0
        a.w_r, a.ldwr, a.w_rm, a.ldwrm = wr.data_ptr(), C, None, 0
        a.peep_i, a.peep_f, a.peep_o = pe[0].data_ptr(), pe[1].data_ptr(), pe[2].data_ptr()
        a.seq_len_dev, a.cell_clip = None, 50.0
        keep.append((wr, pe)); bufs.append(buf)
    wsb = Lb.aslp_lstm_workspace_bytes(T_, S_, C, 0, ndirs, 0)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    K.check(Lb.aslp_lstm_seq_fwd(P(torch.cuda.current_stream().cuda_stream), ctypes.byref(arr), ndirs, P(ws.data_ptr()), wsb))
    torch.cuda.synchronize()
    return [b.cpu().numpy() for b in bufs]


def test_lstm_full_size_stream_permutation_and_direction_mirror():
    C = 320
    rng = np.random.default_rng(9)
    g = (rng.standard_normal((T * S, 4 * C)) * 0.5).astype(np.float32)
    w = (rng.standard_normal((4 * C, C)) * (0.5 / np.sqrt(C))).astype(np.float32)
    peep = [(rng.standard_normal(C) * 0.1).astype(np.float32) for _ in range(3)]
    # direction 1 fed with the time-reversed input of direction 0 and the same weights must produce the time-reversed output
    g3 = g.reshape(T, S, 4 * C)
    base = lstm_run([g, np.ascontiguousarray(g3[::-1]).reshape(T * S, 4 * C)], [w, w], [peep, peep], T, S, C, 2)
    f = base[0][S:(T + 1) * S].reshape(T, S, 7 * C)
    b = base[1][S:(T + 1) * S].reshape(T, S, 7 * C)
    assert np.array_equal(f, b[::-1]), "the two directions are not mirror images"
    # permuting the streams permutes the result bit for bit (streams never mix; the chains are per stream group)
    perm = rng.permutation(S)
    gp = np.ascontiguousarray(g3[:, perm]).reshape(T * S, 4 * C)
    pr = lstm_run([gp], [w], [peep], T, S, C, 1)[0][S:(T + 1) * S].reshape(T, S, 7 * C)
    assert np.array_equal(pr, f[:, perm])
    assert np.isfinite(f).all() and np.abs(f[..., 6 * C:]).max() <= 1.0      # m = tanh(c) * sigmoid(o)


def test_ctc_full_size_gradient_rows_sum_to_zero():
    Lb = K.cuda_lib()
    rng = np.random.default_rng(4)
    acts = torch.from_numpy(rng.standard_normal((T, S, KC)).astype(np.float32)).cuda()
    grads = torch.zeros_like(acts)
    flat = np.ascontiguousarray(rng.integers(1, KC, size=S * L), np.int32)
    llen, ilen = np.full(S, L, np.int32), np.full(S, T, np.int32)
    costs = np.zeros(S, np.float32)
    info = K.CtcComputeInfo(1, torch.cuda.current_stream().cuda_stream)
    size = ctypes.c_size_t(0)
    assert Lb.get_workspace_size(llen.ctypes.data, ilen.ctypes.data, KC, S, info, ctypes.addressof(size)) == 0
    ws = torch.empty(size.value + 256, dtype=torch.uint8, device="cuda")
    assert Lb.compute_ctc_loss(P(acts.data_ptr()), P(grads.data_ptr()), flat.ctypes.data, llen.ctypes.data, ilen.ctypes.data, KC, S,
                               costs.ctypes.data, P(ws.data_ptr()), info) == 0
    g = grads.cpu().numpy().astype(np.float64)
    # grad = softmax - posterior: both sum to one over the classes of a frame
    # (to the accuracy fp32 leaves in exp(. - logZ) at |logZ| ~ 3.7e3: a few ulp of the cost, see the parity test above)
    assert np.abs(g.sum(axis=2)).max() < 64 * float(np.spacing(np.float32(costs.max())))
    assert np.all(costs > 0) and np.all(np.isfinite(costs))
    # -log p(z|x) of random activations is about T * log(K) minus the alignment freedom: sanity window
    assert 2500 < costs.mean() < 4500


@pytest.mark.parametrize("M,N,Kd,ta,tb", [(16000, 1280, 640, 0, 1), (1280, 640, 16000, 1, 0), (16000, 640, 1280, 0, 0)])
def test_gemm_full_size_checksums(M, N, Kd, ta, tb):
    """1^T C = (1^T A) B and C 1 = A (B 1) in float64: O(MK + KN) exact references for a full-size product."""
    Lb = K.cuda_lib()
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((Kd, M) if ta else (M, Kd)).astype(np.float32)
    B = rng.standard_normal((N, Kd) if tb else (Kd, N)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.zeros((M, N), device="cuda")
    wsb = Lb.aslp_gemm_workspace_bytes(M, N, Kd)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    K.check(Lb.aslp_gemm(P(torch.cuda.current_stream().cuda_stream), ta, tb, M, N, Kd, 1.0, P(dA.data_ptr()), A.shape[1], P(dB.data_ptr()),
                         B.shape[1], 0.0, P(dC.data_ptr()), N, P(0), 0.0, 0, P(ws.data_ptr()), wsb))
    C = dC.cpu().numpy().astype(np.float64)
    A64 = (A.T if ta else A).astype(np.float64)
    B64 = (B.T if tb else B).astype(np.float64)
    col = A64.sum(axis=0) @ B64
    row = A64 @ B64.sum(axis=1)
    scale = np.sqrt(Kd) * np.sqrt(M)
    assert np.abs(C.sum(axis=0) - col).max() < 1e-4 * scale
    assert np.abs(C.sum(axis=1) - row).max() < 1e-4 * np.sqrt(Kd) * np.sqrt(N)
