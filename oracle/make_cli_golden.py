"""Trainer-level golden fixtures (tests/golden/cli_*): tiny synthetic ark inputs plus the model and log the UNMODIFIED
reference CPU trainers (oracle/_ref/aslp-nnet-init, aslp-nnet-train-frame, aslp-nnet-train-warp-ctc-streams, built by
oracle/Makefile) write for them with --use-gpu=no.  The GPU tests run OUR binaries (kaldi-aslp_b200/build/bin) with the
same command lines and compare.  Run in the build container: `python oracle/make_cli_golden.py`.  TEST INFRASTRUCTURE ONLY."""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
ENV = dict(os.environ, OPENBLAS_NUM_THREADS="1")


def write_feats_ark(path, utts):
    with open(path, "wb") as f:
        for key, m in utts:
            m = np.ascontiguousarray(m, np.float32)
            f.write(key.encode() + b" \0BFM " + b"\x04" + struct.pack("<i", m.shape[0]) + b"\x04" + struct.pack("<i", m.shape[1]) + m.tobytes())


def write_post_ark(path, utts):
    with open(path, "w") as f:
        for key, ids in utts:
            f.write(key + " " + " ".join("[ %d 1 ]" % i for i in ids) + "\n")


def write_int_ark(path, utts):
    with open(path, "w") as f:
        for key, ids in utts:
            f.write(key + " " + " ".join(str(i) for i in ids) + "\n")


def run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=ENV)
    open(log, "w").write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("%s failed:\n%s" % (" ".join(cmd), r.stdout))


def frame_case():
    d = os.path.join(GOLD, "cli_frame")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(11)
    open(os.path.join(d, "proto.txt"), "w").write(
        "<NnetProto>\n"
        "<AffineTransform> <InputDim> 20 <OutputDim> 32 <BiasMean> -1.0 <BiasRange> 2.0 <ParamStddev> 0.2\n"
        "<Sigmoid> <InputDim> 32 <OutputDim> 32\n"
        "<AffineTransform> <InputDim> 32 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.2\n"
        "<Softmax> <InputDim> 8 <OutputDim> 8\n"
        "</NnetProto>\n")
    feats, post = [], []
    for u in range(7):
        n = int(rng.integers(40, 90))
        feats.append(("utt%02d" % u, rng.standard_normal((n, 20)).astype(np.float32)))
        post.append(("utt%02d" % u, rng.integers(0, 8, size=n).tolist()))
    post = post[:5] + post[6:]                       # utt05 has no targets: must be skipped with a warning
    write_feats_ark(os.path.join(d, "feats.ark"), feats)
    write_post_ark(os.path.join(d, "post.ark"), post)
    run([os.path.join(REF, "aslp-nnet-init"), "--seed=777", "--binary=true", os.path.join(d, "proto.txt"), os.path.join(d, "init.nnet")],
        os.path.join(d, "ref_init.log"))
    args = ["--use-gpu=no", "--minibatch-size=32", "--randomizer-size=200", "--randomizer-seed=777", "--learn-rate=0.02", "--momentum=0.9",
            "--l2-penalty=0.0001", "ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"),
            os.path.join(d, "init.nnet"), os.path.join(d, "ref_out.nnet")]
    run([os.path.join(REF, "aslp-nnet-train-frame")] + args, os.path.join(d, "ref_train.log"))
    open(os.path.join(d, "args.txt"), "w").write("--minibatch-size=32 --randomizer-size=200 --randomizer-seed=777 --learn-rate=0.02 --momentum=0.9 --l2-penalty=0.0001\n")


def frame_mse_case():
    """--objective-function=mse on the cli_frame inputs (Mse through the LossItf pointer of the trainer, nnet-loss.cc:205-290)"""
    src = os.path.join(GOLD, "cli_frame")
    d = os.path.join(GOLD, "cli_frame_mse")
    os.makedirs(d, exist_ok=True)
    flags = "--objective-function=mse --minibatch-size=32 --randomizer-size=200 --randomizer-seed=777 --learn-rate=0.05 --momentum=0.9"
    args = ["--use-gpu=no"] + flags.split() + ["ark:" + os.path.join(src, "feats.ark"), "ark:" + os.path.join(src, "post.ark"),
                                                os.path.join(src, "init.nnet"), os.path.join(d, "ref_out.nnet")]
    run([os.path.join(REF, "aslp-nnet-train-frame")] + args, os.path.join(d, "ref_train.log"))
    open(os.path.join(d, "args.txt"), "w").write(flags + "\n")


def ctc_case():
    d = os.path.join(GOLD, "cli_ctc")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(12)
    open(os.path.join(d, "proto.txt"), "w").write(
        "<NnetProto>\n"
        "<BLstmProjectedStreams> <InputDim> 12 <OutputDim> 16 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2\n"
        "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3\n"
        "<Softmax> <InputDim> 8 <OutputDim> 8\n"
        "</NnetProto>\n")
    feats, labs = [], []
    for u in range(7):
        n = int(rng.integers(20, 40))
        feats.append(("utt%02d" % u, rng.standard_normal((n, 12)).astype(np.float32)))
        labs.append(("utt%02d" % u, rng.integers(1, 8, size=int(rng.integers(3, 8))).tolist()))
    write_feats_ark(os.path.join(d, "feats.ark"), feats)
    write_int_ark(os.path.join(d, "labels.ark"), labs)
    run([os.path.join(REF, "aslp-nnet-init"), "--seed=777", "--binary=true", os.path.join(d, "proto.txt"), os.path.join(d, "init.nnet")],
        os.path.join(d, "ref_init.log"))
    flags = "--num-stream=3 --learn-rate=0.05 --momentum=0.9 --report-period=2 --report-step=2"
    args = ["--use-gpu=no"] + flags.split() + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "labels.ark"),
                                                os.path.join(d, "init.nnet"), os.path.join(d, "ref_out.nnet")]
    run([os.path.join(REF, "aslp-nnet-train-warp-ctc-streams")] + args, os.path.join(d, "ref_train.log"))
    open(os.path.join(d, "args.txt"), "w").write(flags + "\n")


def lc_case():
    d = os.path.join(GOLD, "cli_lc")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(13)
    open(os.path.join(d, "proto.txt"), "w").write(
        "<NnetProto>\n"
        "<BLstmProjectedStreamsLC> <InputDim> 12 <OutputDim> 16 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2\n"
        "<BLstmProjectedStreamsLC> <InputDim> 16 <OutputDim> 16 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2\n"
        "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3\n"
        "<Softmax> <InputDim> 8 <OutputDim> 8\n"
        "</NnetProto>\n")
    feats, post = [], []
    for u in range(8):
        n = int(rng.integers(15, 45))
        feats.append(("utt%02d" % u, rng.standard_normal((n, 12)).astype(np.float32)))
        post.append(("utt%02d" % u, rng.integers(0, 8, size=n).tolist()))
    post[3] = (post[3][0], post[3][1][:-2])          # length mismatch: skipped with a warning, counted as other error
    write_feats_ark(os.path.join(d, "feats.ark"), feats)
    write_post_ark(os.path.join(d, "post.ark"), post)
    run([os.path.join(REF, "aslp-nnet-init"), "--seed=777", "--binary=true", os.path.join(d, "proto.txt"), os.path.join(d, "init.nnet")],
        os.path.join(d, "ref_init.log"))
    flags = "--chunk-size=8 --right-splice=3 --num-stream=3 --learn-rate=0.02 --momentum=0.9 --report-period=3"
    args = ["--use-gpu=no"] + flags.split() + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"),
                                                os.path.join(d, "init.nnet"), os.path.join(d, "ref_out.nnet")]
    run([os.path.join(REF, "aslp-nnet-train-blstm-streams-lc")] + args, os.path.join(d, "ref_train.log"))
    open(os.path.join(d, "args.txt"), "w").write(flags + "\n")


def lstm_case():
    d = os.path.join(GOLD, "cli_lstm")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(14)
    open(os.path.join(d, "proto.txt"), "w").write(
        "<NnetProto>\n"
        "<Lstm> <InputDim> 12 <OutputDim> 16 <ClipGradient> 5 <ParamScale> 0.2\n"
        "<LstmProjectedStreams> <InputDim> 16 <OutputDim> 8 <CellDim> 16 <ClipGradient> 5 <ParamScale> 0.2\n"
        "<AffineTransform> <InputDim> 8 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3\n"
        "<Softmax> <InputDim> 8 <OutputDim> 8\n"
        "</NnetProto>\n")
    feats, post = [], []
    for u in range(7):
        n = int(rng.integers(9, 30))
        feats.append(("utt%02d" % u, rng.standard_normal((n, 12)).astype(np.float32)))
        post.append(("utt%02d" % u, rng.integers(0, 8, size=n).tolist()))
    write_feats_ark(os.path.join(d, "feats.ark"), feats)
    write_post_ark(os.path.join(d, "post.ark"), post)
    run([os.path.join(REF, "aslp-nnet-init"), "--seed=777", "--binary=true", os.path.join(d, "proto.txt"), os.path.join(d, "init.nnet")],
        os.path.join(d, "ref_init.log"))
    flags = "--batch-size=5 --num-stream=3 --targets-delay=2 --learn-rate=0.02 --momentum=0.9 --report-period=2"
    args = ["--use-gpu=no"] + flags.split() + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"),
                                                os.path.join(d, "init.nnet"), os.path.join(d, "ref_out.nnet")]
    run([os.path.join(REF, "aslp-nnet-train-lstm-streams")] + args, os.path.join(d, "ref_train.log"))
    open(os.path.join(d, "args.txt"), "w").write(flags + "\n")


def fsmn_case():
    d = os.path.join(GOLD, "cli_fsmn")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(15)
    open(os.path.join(d, "proto.txt"), "w").write(
        "<NnetProto>\n"
        "<AffineTransform> <InputDim> 12 <OutputDim> 24 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.2\n"
        "<ReLU> <InputDim> 24 <OutputDim> 24\n"
        "<AffineTransform> <InputDim> 24 <OutputDim> 16 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.2\n"
        "<CompactFsmn> <InputDim> 16 <OutputDim> 16 <PastContext> 4 <FutureContext> 3\n"
        "<AffineTransform> <InputDim> 16 <OutputDim> 24 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.2\n"
        "<ReLU> <InputDim> 24 <OutputDim> 24\n"
        "<AffineTransform> <InputDim> 24 <OutputDim> 16 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.2\n"
        "<CompactFsmn> <InputDim> 16 <OutputDim> 16 <PastContext> 4 <FutureContext> 3\n"
        "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3\n"
        "<Softmax> <InputDim> 8 <OutputDim> 8\n"
        "</NnetProto>\n")
    feats, post = [], []
    for u in range(6):
        n = int(rng.integers(20, 60))
        feats.append(("utt%02d" % u, rng.standard_normal((n, 12)).astype(np.float32)))
        post.append(("utt%02d" % u, rng.integers(0, 8, size=n).tolist()))
    post[2] = (post[2][0], post[2][1][:-3])          # 3 targets short: inside --length-tolerance=5, truncated to the minimum
    post[4] = (post[4][0], post[4][1][:-9])          # 9 short: skipped, counted as other error
    write_feats_ark(os.path.join(d, "feats.ark"), feats)
    write_post_ark(os.path.join(d, "post.ark"), post)
    run([os.path.join(REF, "aslp-nnet-init"), "--seed=777", "--binary=true", os.path.join(d, "proto.txt"), os.path.join(d, "init.nnet")],
        os.path.join(d, "ref_init.log"))
    flags = "--learn-rate=20.0 --momentum=0.5 --report-period=60"
    args = ["--use-gpu=no"] + flags.split() + ["ark:" + os.path.join(d, "feats.ark"), "ark:" + os.path.join(d, "post.ark"),
                                                os.path.join(d, "init.nnet"), os.path.join(d, "ref_out.nnet")]
    run([os.path.join(REF, "aslp-nnet-train-perutt")] + args, os.path.join(d, "ref_train.log"))
    open(os.path.join(d, "args.txt"), "w").write(flags + "\n")


def forward_cases():
    """Forwarders: the unmodified reference aslp-nnet-forward / aslp-nnet-forward-blstm-lc / aslp-nnet-forward-skip on the models the trainer cases
    above wrote.  Each entry: (name, binary, model case, flags)."""
    d = os.path.join(GOLD, "cli_fwd")
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "counts.txt"), "w").write(" [ 30 10 0 25 5 40 20 15 ]\n")     # one empty class: floored prior
    cases = [
        ("frame_default", "aslp-nnet-forward", "cli_frame", ""),
        ("frame_prior", "aslp-nnet-forward", "cli_frame", "--class-frame-counts=%s --prior-scale=0.8" % os.path.join(d, "counts.txt")),
        ("frame_skip", "aslp-nnet-forward", "cli_frame", "--skip-width=3 --apply-log=false"),
        ("lstm_shift", "aslp-nnet-forward", "cli_lstm", "--time-shift=2 --apply-log=false"),
        ("ctc_blank", "aslp-nnet-forward", "cli_ctc", "--add-softmax=true --scale-blank=0.5"),
        ("lc_chunks", "aslp-nnet-forward-blstm-lc", "cli_lc", "--chunk-size=8 --right-splice=3"),
        ("skip_split3", "aslp-nnet-forward-skip", "cli_frame", "--skip-width=3"),
        ("skip_lstm2", "aslp-nnet-forward-skip", "cli_lstm", "--skip-width=2 --apply-log=false --time-shift=1"),
        ("skip_ctc4", "aslp-nnet-forward-skip", "cli_ctc", "--skip-width=4 --add-softmax=true --scale-blank=0.25"),
        ("lc_prior", "aslp-nnet-forward-blstm-lc", "cli_lc", "--chunk-size=8 --right-splice=3 --apply-log=false --no-softmax=true "
         "--class-frame-counts=%s" % os.path.join(d, "counts.txt")),
    ]
    with open(os.path.join(d, "cases.txt"), "w") as f:
        for name, exe, case, flags in cases:
            src = os.path.join(GOLD, case)
            out = os.path.join(d, name + ".ark")
            run([os.path.join(REF, exe), "--use-gpu=no"] + flags.split() + [os.path.join(src, "ref_out.nnet"), "ark:" + os.path.join(src, "feats.ark"), "ark:" + out],
                os.path.join(d, name + ".log"))
            f.write("%s\t%s\t%s\t%s\n" % (name, exe, case, flags.replace(d + os.sep, "@GOLD@/")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "forward":
        forward_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "frame_mse":
        frame_mse_case()
        sys.exit(0)
    frame_case()
    frame_mse_case()
    ctc_case()
    lc_case()
    lstm_case()
    fsmn_case()
    forward_cases()
    for c in ("cli_frame", "cli_ctc", "cli_lc", "cli_lstm", "cli_fsmn"):
        d = os.path.join(GOLD, c)
        print(c, sorted(os.listdir(d)), sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)), "bytes")
