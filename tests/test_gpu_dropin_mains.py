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
