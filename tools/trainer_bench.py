"""End-to-end rate of the trainer MAIN (data feeding): aslp-nnet-train-warp-ctc-streams on a synthetic
Kaldi archive of BASELINE cfg3 shape (utterances of 1000 frames x 40 dims, 100 labels of 72 classes, 16 streams), read from
disk, with the feeder thread (ASLP_FEEDER_DEPTH=2, page-locked slots) and without (0: the reference's inline read / pack /
copy order).  Prints the trainer's own fps figure and the wall time of the training loop for both."""
import json
import os
import re
import struct
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "kaldi-aslp_b200", "build", "bin")
PROTO = """<NnetProto>
<BLstmProjectedStreamsLC> <InputDim> 40 <OutputDim> 640 <CellDim> 320 <ParamScale> 0.01 <ClipGradient> 5
<BLstmProjectedStreamsLC> <InputDim> 640 <OutputDim> 640 <CellDim> 320 <ParamScale> 0.01 <ClipGradient> 5
<BLstmProjectedStreamsLC> <InputDim> 640 <OutputDim> 640 <CellDim> 320 <ParamScale> 0.01 <ClipGradient> 5
<AffineTransform> <InputDim> 640 <OutputDim> 72 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.05
<Softmax> <InputDim> 72 <OutputDim> 72
</NnetProto>
"""


def write_archives(td, utts, frames, dim, labels, classes, seed=1):
    rng = np.random.default_rng(seed)
    with open(os.path.join(td, "feats.ark"), "wb") as f, open(os.path.join(td, "labels.ark"), "w") as g:
        for u in range(utts):
            key = ("utt%05d " % u).encode()
            m = rng.standard_normal((frames, dim), dtype=np.float32)
            f.write(key + b"\0BFM " + b"\4" + struct.pack("<i", frames) + b"\4" + struct.pack("<i", dim) + m.tobytes())
            lab = rng.integers(1, classes, labels, dtype=np.int32)
            g.write(key.decode() + " ".join(str(int(v)) for v in lab) + "\n")       # text int32-vector archive


def main():
    utts = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    with tempfile.TemporaryDirectory() as td:
        write_archives(td, utts, 1000, 40, 100, 72)
        open(os.path.join(td, "proto.txt"), "w").write(PROTO)
        subprocess.check_call([os.path.join(BIN, "aslp-nnet-init"), "--seed=777", "--binary=true", os.path.join(td, "proto.txt"),
                               os.path.join(td, "init.nnet")], stderr=subprocess.DEVNULL)
        out = {}
        for rep in range(3):
            for depth in ["0", "2"]:
                env = dict(os.environ, ASLP_FEEDER_DEPTH=depth)
                t0 = time.time()
                r = subprocess.run([os.path.join(BIN, "aslp-nnet-train-warp-ctc-streams"), "--num-stream=16", "--frame-limit=1000000",
                                    "--learn-rate=0.016", "--momentum=0.9", "--report-period=100000",
                                    "ark:" + os.path.join(td, "feats.ark"), "ark:" + os.path.join(td, "labels.ark"),
                                    os.path.join(td, "init.nnet"), os.path.join(td, "out%s.nnet" % depth)],
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
                wall = time.time() - t0
                assert r.returncode == 0, r.stdout[-2000:]
                fps = float(re.findall(r"fps([\d.eE+-]+)\]", r.stdout)[-1])
                out.setdefault("feeder_depth_" + depth, []).append({"trainer_fps": fps, "process_wall_s": round(wall, 3)})
        same = open(os.path.join(td, "out0.nnet"), "rb").read() == open(os.path.join(td, "out2.nnet"), "rb").read()
        print(json.dumps({"what": "aslp-nnet-train-warp-ctc-streams, cfg3 net, %d utterances x 1000 frames from an ark on local disk, 16 streams" % utts,
                          "identical_models": same, **out}))


if __name__ == "__main__":
    main()
