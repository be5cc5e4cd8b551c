"""Drop-in at run time: the UNMODIFIED reference trainer mains (src/aslp-nnetbin/*.cc, src/aslp-parallelbin/*.cc), compiled by
`make -C oracle dropin` against OUR host layer and kernels, run on the GPU on the trainer-level golden fixtures and must write the
model the reference's own CPU build wrote for the same command line (tests/golden/cli_*, 1e-4).  What runs is the reference's
main() -- its option parsing, its batching loops, its calls into Nnet / Xent / WarpCtc / IWorker -- on top of libaslp_nnet.so."""
import os
import re
import subprocess

import numpy as np
import pytest

from tests.test_gpu_cli import params, GOLD, RTOL

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")

needs_dropin = pytest.mark.skipif(not os.path.exists(os.path.join(DROPIN, "aslp-nnet-train-frame")), reason="oracle/_ref/dropin not built")


def run(exe, args, env=None):
    r = subprocess.run([os.path.join(DROPIN, exe)] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def check_model(out, case):
    got, want = params(out), params(os.path.join(GOLD, case, "ref_out.nnet"))
    assert got.shape == want.shape
    err = np.max(np.abs(got - want)) / np.max(np.abs(want))
    assert err < RTOL, (case, err)


@needs_dropin
@pytest.mark.parametrize("exe,case,targets", [
    ("aslp-nnet-train-frame", "cli_frame", "post.ark"),
    ("aslp-nnet-train-warp-ctc-streams", "cli_ctc", "labels.ark"),
    ("aslp-nnet-train-blstm-streams-lc", "cli_lc", "post.ark"),
    ("aslp-nnet-train-lstm-streams", "cli_lstm", "post.ark"),
])
def test_unmodified_reference_main_on_our_library(exe, case, targets, tmp_path):
    d = os.path.join(GOLD, case)
    out = str(tmp_path / "out.nnet")
    flags = open(os.path.join(d, "args.txt")).read().split()
    log = run(exe, flags + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, targets), os.path.join(d, "init.nnet"), out])
    ref_log = open(os.path.join(d, "ref_train.log")).read()
    done = lambda s: re.findall(r"Done (\d+) files", s)[-1:]          # the frame trainer reports frames, not files
    assert done(log) == done(ref_log)
    frames = lambda s: re.findall(r"Frame: (\d+)", s)[-1:]
    assert frames(log) == frames(ref_log)
    check_model(out, case)


@needs_dropin
def test_unmodified_reference_worker_main_single_rank(tmp_path):
    """the reference's LC-BLSTM worker main (new BmufWorker(momentum, learn_rate) bootstrapped from the environment) as the only
    rank: block momentum 0 and block learn rate 1 make BMUF the identity filter, so it must reproduce the plain trainer's model"""
    d = os.path.join(GOLD, "cli_lc")
    out = str(tmp_path / "out.nnet")
    flags = [f.replace("--right-splice", "--right_splice") for f in open(os.path.join(d, "args.txt")).read().split()]
    flags += ["--worker-type=bmuf", "--bmuf-momentum=0", "--bmuf-learn-rate=1", "--sync-period=40"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", ASLP_NCCL_ID_FILE=str(tmp_path / "nccl_id"))
    run("aslp-nnet-train-lc-blstm-streams-worker",
        flags + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"), os.path.join(d, "init.nnet"), out], env)
    check_model(out, "cli_lc")


@needs_dropin
def test_unmodified_reference_info_main():
    log = run("aslp-nnet-info", [os.path.join(GOLD, "cli_lc", "init.nnet")])
    assert "BLstmProjectedStreamsLC" in log and "num-components" in log


@needs_dropin
def test_unmodified_reference_perutt_main_on_our_library(tmp_path):
    """aslp-nnet-train-perutt (BASELINE cfg4's trainer): frame-weight / utterance-weight table readers, length tolerance, the
    learn_rate / 1024 quirk -- the reference's own main() on our classes, against the model its CPU build wrote"""
    d = os.path.join(GOLD, "cli_fsmn")
    out = str(tmp_path / "out.nnet")
    flags = open(os.path.join(d, "args.txt")).read().split()
    log = run("aslp-nnet-train-perutt", flags + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"),
                                                 os.path.join(d, "init.nnet"), out])
    ref_log = open(os.path.join(d, "ref_train.log")).read()
    done = lambda s: re.findall(r"Done (\d+) files, (\d+) with no tgt_mats, (\d+) with other errors", s)[-1]
    assert done(log) == done(ref_log)
    check_model(out, "cli_fsmn")


def _fwd_cases():
    from tests.test_gpu_cli import fwd_cases
    return fwd_cases()


@needs_dropin
@pytest.mark.parametrize("name,exe,case,flags", _fwd_cases(), ids=[c[0] for c in _fwd_cases()])
def test_unmodified_reference_forwarders_on_our_library(name, exe, case, flags, tmp_path):
    """the reference's forwarder mains (CuMatrix::Row / Min / Max / Add / ApplyLog / ApplySoftMaxPerRow, PdfPrior::SubtractOnLogpost,
    Matrix::RowRange ...) on our library, against the archives their own CPU build wrote (tests/golden/cli_fwd)"""
    from tests.test_gpu_cli import read_feats_ark
    if not os.path.exists(os.path.join(DROPIN, exe)):
        pytest.skip(exe + " not built")
    d = os.path.join(GOLD, "cli_fwd")
    out = str(tmp_path / "out.ark")
    args = flags.replace("@GOLD@", d).split()
    run(exe, args + [os.path.join(GOLD, case, "ref_out.nnet"), "ark:" + os.path.join(GOLD, case, "feats.ark"), "ark:" + out])
    want, got = read_feats_ark(os.path.join(d, name + ".ark")), read_feats_ark(out)
    assert list(got) == list(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        big = np.abs(want[k]) >= 1e18
        assert np.array_equal(big, np.abs(got[k]) >= 1e18), k
        scale = max(1.0, float(np.max(np.abs(want[k][~big]))))
        assert np.max(np.abs(got[k][~big] - want[k][~big])) / scale < RTOL, (k, np.max(np.abs(got[k][~big] - want[k][~big])))


@needs_dropin
def test_unmodified_reference_init_and_copy_mains(tmp_path):
    """aslp-nnet-init --seed=777 must write the reference's initial model; aslp-nnet-copy must round-trip it byte for byte"""
    d = os.path.join(GOLD, "cli_frame")
    if not os.path.exists(os.path.join(d, "proto.txt")) or not os.path.exists(os.path.join(DROPIN, "aslp-nnet-init")):
        pytest.skip("fixture or binary missing")
    out = str(tmp_path / "init.nnet")
    run("aslp-nnet-init", ["--seed=777", "--binary=true", os.path.join(d, "proto.txt"), out])
    got, want = params(out), params(os.path.join(d, "init.nnet"))
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-6 * np.max(np.abs(want))
    cp = str(tmp_path / "copy.nnet")
    run("aslp-nnet-copy", ["--binary=true", os.path.join(d, "init.nnet"), cp])
    assert open(cp, "rb").read() == open(os.path.join(d, "init.nnet"), "rb").read()


@needs_dropin
def test_unmodified_reference_mimo_frame_trainer_with_one_stream(tmp_path):
    """aslp-nnet-train-frame-mimo (the multi-stream FrameDataReader, vector Propagate / Backpropagate, one loss per output) on a
    net with one input and one output: the same shuffle mask, minibatches and updates as the plain frame trainer, so it must
    write the model the reference's aslp-nnet-train-frame wrote for the fixture"""
    if not os.path.exists(os.path.join(DROPIN, "aslp-nnet-train-frame-mimo")):
        pytest.skip("aslp-nnet-train-frame-mimo not built")
    d = os.path.join(GOLD, "cli_frame")
    out = str(tmp_path / "out.nnet")
    flags = open(os.path.join(d, "args.txt")).read().split()
    run("aslp-nnet-train-frame-mimo", flags + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"),
                                               os.path.join(d, "init.nnet"), out])
    check_model(out, "cli_frame")
