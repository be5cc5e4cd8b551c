"""Generates the committed golden fixtures under tests/golden/ by running the UNMODIFIED reference CPU build
(oracle/_ref/ref_driver, see oracle/Makefile) on small seeded cases.  Run in the build container (needs /root/reference
only to have built oracle/_ref): `python oracle/make_golden.py`.  TEST INFRASTRUCTURE ONLY."""
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import kaldi_io  # noqa: E402

DRIVER = os.path.join(HERE, "_ref", "ref_driver")
GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (proto lines, input dim, rows, spec dict, loss)
CASES = {
    "dnn_xent": dict(
        proto=["<AffineTransform> <InputDim> 20 <OutputDim> 32 <BiasMean> -1.0 <BiasRange> 2.0 <ParamStddev> 0.2",
               "<Sigmoid> <InputDim> 32 <OutputDim> 32",
               "<AffineTransform> <InputDim> 32 <OutputDim> 16 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.2 <LearnRateCoef> 0.5 <BiasLearnRateCoef> 2.0",
               "<Softmax> <InputDim> 16 <OutputDim> 16"],
        dim=20, rows=24, loss="xent", spec=dict(learn_rate=0.1, momentum=0.9, iters=2, dump_components=1)),
    "dnn_l1l2_maxnorm": dict(
        proto=["<AffineTransform> <InputDim> 12 <OutputDim> 20 <BiasMean> 0 <BiasRange> 1.0 <ParamStddev> 0.3 <MaxNorm> 0.8",
               "<Tanh> <InputDim> 20 <OutputDim> 20",
               "<LinearTransform> <InputDim> 20 <OutputDim> 20 <ParamStddev> 0.2",
               "<ReLU> <InputDim> 20 <OutputDim> 20",
               "<AffineTransform> <InputDim> 20 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.2",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=12, rows=16, loss="xent", spec=dict(learn_rate=0.05, momentum=0.5, l2=0.001, l1=0.0005, iters=2)),
    "cnn_xent": dict(           # Conv (8 patch positions x 8 filters, dense single-GEMM view) + overlapping max pooling
        proto=["<ConvolutionalComponent> <InputDim> 33 <OutputDim> 64 <PatchDim> 4 <PatchStep> 1 <PatchStride> 11 <BiasMean> -0.5 <BiasRange> 1.0 <ParamStddev> 0.3",
               "<MaxPoolingComponent> <InputDim> 64 <OutputDim> 24 <PoolSize> 4 <PoolStep> 2 <PoolStride> 8",
               "<Sigmoid> <InputDim> 24 <OutputDim> 24",
               "<AffineTransform> <InputDim> 24 <OutputDim> 5 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 5 <OutputDim> 5"],
        dim=33, rows=20, loss="xent", spec=dict(learn_rate=0.1, momentum=0.9, iters=2, dump_components=1)),
    "cnn_odd_maxnorm": dict(    # 7 x 5 = 35 output columns (not a multiple of 4: the per-patch form), max-norm, lr coefficients
        proto=["<ConvolutionalComponent> <InputDim> 30 <OutputDim> 35 <PatchDim> 4 <PatchStep> 1 <PatchStride> 10 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.4 "
               "<LearnRateCoef> 0.7 <BiasLearnRateCoef> 1.5 <MaxNorm> 0.9",
               "<MaxPoolingComponent> <InputDim> 35 <OutputDim> 15 <PoolSize> 3 <PoolStep> 2 <PoolStride> 5",
               "<Tanh> <InputDim> 15 <OutputDim> 15",
               "<AffineTransform> <InputDim> 15 <OutputDim> 4 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 4 <OutputDim> 4"],
        dim=30, rows=13, loss="xent", spec=dict(learn_rate=0.2, momentum=0.5, iters=3, dump_components=1)),
    "lstm_xent": dict(
        proto=["<Lstm> <InputDim> 10 <OutputDim> 16 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=6 * 3, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=3, reset_flags="1,0,1", dump_components=1)),
    "lstmp_xent": dict(
        proto=["<LstmProjectedStreams> <InputDim> 10 <OutputDim> 12 <CellDim> 16 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 12 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=5 * 4, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, reset_flags="0,0,0,0")),
    "blstm_seqlen": dict(
        proto=["<BLstm> <InputDim> 10 <OutputDim> 16 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=6 * 3, loss="none", spec=dict(learn_rate=0.05, momentum=0.0, iters=1, seq_lengths="6,4,5", dump_components=1)),
    "blstmp_seqlen": dict(
        proto=["<BLstmProjectedStreams> <InputDim> 10 <OutputDim> 12 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 12 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=6 * 3, loss="none", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, seq_lengths="6,3,5")),
    "lc_blstm_chunks": dict(
        proto=["<BLstmProjectedStreamsLC> <InputDim> 10 <OutputDim> 12 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2",
               "<BLstmProjectedStreamsLC> <InputDim> 12 <OutputDim> 12 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 12 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=6 * 4, loss="xent", mask_tail=(4, 6, 4),     # chunk 4 + right splice 2, S = 4: mask = 1 only for t < 4
        spec=dict(learn_rate=0.05, momentum=0.9, iters=3, chunk_size=4, reset_flags="1,0,0,1", dump_components=1)),
    "lc_blstm_ctc": dict(
        proto=["<BLstmProjectedStreamsLC> <InputDim> 10 <OutputDim> 16 <CellDim> 8 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.4",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=12 * 3, loss="ctc", labels=[[1, 2, 2, 3], [4, 5], [6, 1, 7]],
        spec=dict(momentum=0.9, iters=2, seq_lengths="12,12,12", norm_learn_rate=0.5, dump_components=1)),
    "gru_xent": dict(
        proto=["<GruStreams> <InputDim> 10 <OutputDim> 12 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 12 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=5 * 3, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, reset_flags="1,1,1", dump_components=1)),
    "fsmn_xent": dict(
        proto=["<AffineTransform> <InputDim> 12 <OutputDim> 24 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.2",
               "<ReLU> <InputDim> 24 <OutputDim> 24",
               "<AffineTransform> <InputDim> 24 <OutputDim> 16 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.2",
               "<CompactFsmn> <InputDim> 16 <OutputDim> 16 <PastContext> 5 <FutureContext> 3",
               "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=12, rows=30, loss="xent", spec=dict(learn_rate=0.02, momentum=0.0, iters=2, dump_components=1)),
    "bn_xent": dict(
        proto=["<AffineTransform> <InputDim> 12 <OutputDim> 20 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.3",
               "<BatchNormalization> <InputDim> 20 <OutputDim> 20",
               "<Sigmoid> <InputDim> 20 <OutputDim> 20",
               "<AffineTransform> <InputDim> 20 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=12, rows=32, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, dump_components=1)),
    "zoo_pnorm_lengthnorm": dict(      # Pnorm (p = 2 and a general p), LengthNorm, Copy as a front end
        proto=["<Copy> <InputDim> 10 <OutputDim> 12 <BuildVector> 1:10 3 7 </BuildVector>",
               "<AffineTransform> <InputDim> 12 <OutputDim> 24 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.3",
               "<Pnorm> <InputDim> 24 <OutputDim> 8 <P> 2",
               "<LengthNormComponent> <InputDim> 8 <OutputDim> 8",
               "<AffineTransform> <InputDim> 8 <OutputDim> 18 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.4",
               "<Pnorm> <InputDim> 18 <OutputDim> 6 <P> 3",
               "<AffineTransform> <InputDim> 6 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=21, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, dump_components=1)),
    "zoo_blocksoftmax_mse": dict(      # BlockSoftmax over blocks of 5 + 3 outputs, trained with the Mse objective
        proto=["<AffineTransform> <InputDim> 9 <OutputDim> 14 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.3",
               "<Tanh> <InputDim> 14 <OutputDim> 14",
               "<AffineTransform> <InputDim> 14 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<BlockSoftmax> <InputDim> 8 <OutputDim> 8 <BlockDims> 5:3"],
        dim=9, rows=17, loss="mse", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, dump_components=1)),
    "zoo_blocksoftmax_xent": dict(     # the backward rule that zeroes the block without the target
        proto=["<AffineTransform> <InputDim> 9 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.3",
               "<BlockSoftmax> <InputDim> 8 <OutputDim> 8 <BlockDims> 5:3"],
        dim=9, rows=16, loss="xent", spec=dict(learn_rate=0.05, momentum=0.5, iters=2, dump_components=1)),
    "zoo_dropout": dict(               # masks drawn from rand() in the reference CPU order (ASLP_DROPOUT_HOST_RAND=1 on our side)
        proto=["<AffineTransform> <InputDim> 10 <OutputDim> 16 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.3",
               "<Sigmoid> <InputDim> 16 <OutputDim> 16",
               "<Dropout> <InputDim> 16 <OutputDim> 16",                       # the reference cannot parse <DropoutRetention> from a proto line (ReadToken fails at the line end): default 0.5
               "<AffineTransform> <InputDim> 16 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=19, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=3, srand=4321, dump_components=1)),
    "lstm_cifg_xent": dict(            # coupled input-forget gate, state carried over the iterations, one stream restarted
        proto=["<LstmCifgProjectedStreams> <InputDim> 10 <OutputDim> 12 <CellDim> 16 <ClipGradient> 5 <ParamScale> 0.2",
               "<AffineTransform> <InputDim> 12 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=10, rows=6 * 3, loss="xent", spec=dict(learn_rate=0.05, momentum=0.9, iters=3, reset_flags="1,0,0", dump_components=1)),
    "splice_rowconv": dict(
        proto=["<Splice> <InputDim> 6 <OutputDim> 18 <BuildVector> -1:1 </BuildVector>",
               "<AffineTransform> <InputDim> 18 <OutputDim> 12 <BiasMean> 0 <BiasRange> 0.5 <ParamStddev> 0.3",
               "<RowConvolution> <InputDim> 12 <OutputDim> 12 <FutureContext> 2",
               "<AffineTransform> <InputDim> 12 <OutputDim> 8 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.3",
               "<Softmax> <InputDim> 8 <OutputDim> 8"],
        dim=6, rows=7 * 2, loss="none", spec=dict(learn_rate=0.05, momentum=0.9, iters=2, seq_lengths="7,4", dump_components=1)),
}


def main():
    if not os.path.exists(DRIVER):
        raise SystemExit("build oracle/_ref first: make -C oracle")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    only = sys.argv[1:]                       # optional: regenerate just the named cases
    for name, c in CASES.items():
        if only and name not in only:
            continue
        d = os.path.join(GOLD, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        with open(os.path.join(d, "proto.txt"), "w") as f:
            f.write("<NnetProto>\n" + "\n".join(c["proto"]) + "\n</NnetProto>\n")
        subprocess.check_call([DRIVER, "init", os.path.join(d, "proto.txt"), os.path.join(d, "model.bin"), "777", "1"], env=env,
                              stderr=subprocess.DEVNULL)
        x = rng.standard_normal((c["rows"], c["dim"])).astype(np.float32)
        kaldi_io.write_mat(os.path.join(d, "input.mat"), x)
        out_dim = int(c["proto"][-1].split("<OutputDim>")[1].split()[0])
        spec = dict(c["spec"])
        spec["input"] = "input.mat"
        spec["loss"] = c["loss"]
        if c["loss"] in ("xent", "mse"):
            np.savetxt(os.path.join(d, "targets.txt"), rng.integers(0, out_dim, c["rows"]), fmt="%d")
            spec["targets"] = "targets.txt"
            if "mask_tail" in c:
                chunk, T, S = c["mask_tail"]
                mask = np.array([1.0 if t < chunk else 0.0 for t in range(T) for _ in range(S)], np.float32)
                np.savetxt(os.path.join(d, "frame_mask.txt"), mask, fmt="%g")
                spec["frame_mask"] = "frame_mask.txt"
        elif c["loss"] == "ctc":
            with open(os.path.join(d, "labels.txt"), "w") as f:
                for l in c["labels"]:
                    f.write(" ".join(map(str, l)) + "\n")
            spec["labels"] = "labels.txt"
        else:
            kaldi_io.write_mat(os.path.join(d, "out_diff.mat"), (rng.standard_normal((c["rows"], out_dim)) * 0.1).astype(np.float32))
            spec["out_diff"] = "out_diff.mat"
        with open(os.path.join(d, "spec.txt"), "w") as f:
            for k, v in spec.items():
                f.write("%s %s\n" % (k, v))
        subprocess.check_call([DRIVER, "step", "model.bin", "spec.txt", "ref"], cwd=d, env=env, stderr=subprocess.DEVNULL)
        size = sum(os.path.getsize(os.path.join(dp, fn)) for dp, _, fns in os.walk(d) for fn in fns)
        print("%-20s %6.1f KB" % (name, size / 1024))


if __name__ == "__main__":
    main()
