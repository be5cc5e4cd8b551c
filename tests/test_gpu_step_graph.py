"""Static-shape step replay (host/nnet-train-step.h, CuStepGraph): a training step that is recorded once and replayed must
leave exactly the parameters, outputs and loss statistics the enqueued step leaves -- same kernels, same order, so the
comparison is bit for bit.  Each arm runs in its own process (ASLP_STEP_GRAPH is read once per process)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, %(root)r)
from kaldi_aslp_b200 import nnet as NN
kind, out = sys.argv[1], sys.argv[2]
protos = {
  "dnn": "<AffineTransform> <InputDim> 120 <OutputDim> 256 <BiasMean> -2.0 <BiasRange> 4.0 <ParamStddev> 0.04\n<Sigmoid> <InputDim> 256 <OutputDim> 256\n"
         "<AffineTransform> <InputDim> 256 <OutputDim> 256 <BiasMean> -2.0 <BiasRange> 4.0 <ParamStddev> 0.04 <MaxNorm> 2.0\n<Tanh> <InputDim> 256 <OutputDim> 256\n"
         "<AffineTransform> <InputDim> 256 <OutputDim> 300 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04\n<Softmax> <InputDim> 300 <OutputDim> 300\n",
  "fsmn": "<AffineTransform> <InputDim> 120 <OutputDim> 256 <BiasMean> 0 <BiasRange> 0.1 <ParamStddev> 0.04\n<ReLU> <InputDim> 256 <OutputDim> 256\n"
          "<LinearTransform> <InputDim> 256 <OutputDim> 128 <ParamStddev> 0.04\n<CompactFsmn> <InputDim> 128 <OutputDim> 128 <PastContext> 5 <FutureContext> 3\n"
          "<AffineTransform> <InputDim> 128 <OutputDim> 300 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04\n<Softmax> <InputDim> 300 <OutputDim> 300\n",
}
NN.select_device(0)
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, "proto.txt")
    open(p, "w").write("<NnetProto>\n" + protos[kind] + "</NnetProto>\n")
    NN.srand(777)
    net = NN.Nnet.init(p)
net.set_train_options(learn_rate=1e-3, momentum=0.9, l2_penalty=1e-5, l1_penalty=1e-6)
rng = np.random.default_rng(3)
xent = NN.Xent()
# two interleaved minibatch shapes (as a frame trainer's tail or a per-utterance trainer produce), a masked frame, a changed learn rate
shapes = [256] * 5 + [192] * 4 + [256] * 2 + [256] * 4
for step, rows in enumerate(shapes):
    if step == 11:
        net.set_train_options(learn_rate=5e-4, momentum=0.9, l2_penalty=1e-5, l1_penalty=1e-6)
    x = rng.standard_normal((rows, 120)).astype(np.float32)
    t = rng.integers(0, 300, size=rows).astype(np.int32)
    mask = np.ones(rows, np.float32); mask[rows // 3] = 0.0
    NN.train_step_xent(net, xent, x, t, frame_mask=mask)
np.save(out + ".params.npy", net.get_params())
np.save(out + ".out.npy", net.component_output(net.num_components - 1, shapes[-1], 300))
json.dump({"report": xent.report(), "replays": NN.step_replays()}, open(out + ".json", "w"))
'''


@pytest.mark.parametrize("kind", ["dnn", "fsmn"])
def test_replayed_steps_are_bit_identical_to_enqueued_steps(kind, tmp_path):
    script = tmp_path / "run.py"
    script.write_text(SCRIPT % {"root": ROOT})
    res = {}
    for arm, env in (("graph", "1"), ("eager", "0")):
        e = dict(os.environ, ASLP_STEP_GRAPH=env)
        out = str(tmp_path / arm)
        r = subprocess.run([sys.executable, str(script), kind, out], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[arm] = (np.load(out + ".params.npy"), np.load(out + ".out.npy"), json.load(open(out + ".json")))
    assert res["eager"][2]["replays"] == 0
    assert res["graph"][2]["replays"] >= 5, res["graph"][2]          # both shapes and both learn rates were recorded and replayed
    assert np.array_equal(res["graph"][0], res["eager"][0])           # parameters after 15 steps: bit for bit
    assert np.array_equal(res["graph"][1], res["eager"][1])           # last output
    assert res["graph"][2]["report"] == res["eager"][2]["report"]     # loss / accuracy bookkeeping


@pytest.mark.parametrize("kind", ["dnn", "fsmn"])
def test_epilogue_fusion_matches_the_unfused_component_sequence(kind, tmp_path):
    """Nnet pairs Affine + activation (forward) and activation + Affine (backward) into one product each and folds the SGD
    apply into the weight-gradient product (aslp_gemm_ex); ASLP_FUSE_EPILOGUE=0 runs the components one by one.  Same
    arithmetic per element -- only the summation order of the bias gradient differs -- so 15 steps agree to 1e-5."""
    script = tmp_path / "run.py"
    script.write_text(SCRIPT % {"root": ROOT})
    res = {}
    for arm, env in (("fused", "1"), ("unfused", "0")):
        e = dict(os.environ, ASLP_STEP_GRAPH="0", ASLP_FUSE_EPILOGUE=env)
        out = str(tmp_path / arm)
        r = subprocess.run([sys.executable, str(script), kind, out], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[arm] = (np.load(out + ".params.npy"), np.load(out + ".out.npy"))
    for a, b in zip(res["fused"], res["unfused"]):
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
