"""Device-resident training-step throughput of the BASELINE configs that bench.py does not time (cfg1 DNN, cfg2 LSTM, cfg4 FSMN;
cfg3 is the bench line): one minibatch = Propagate + Xent + Backpropagate with all updates through the handle API, random-init
weights of the named architecture, synthetic features, CUDA events, median of 7 after 3 warm-ups.  Informative only."""
import ctypes
import json
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from kaldi_aslp_b200 import nnet as NN  # noqa: E402

CFG = {
    "cfg1 DNN 440-4x1024-1500, minibatch 256": dict(
        proto="".join(["<AffineTransform> <InputDim> %d <OutputDim> 1024 <BiasMean> -2.0 <BiasRange> 4.0 <ParamStddev> 0.04\n<Sigmoid> <InputDim> 1024 <OutputDim> 1024\n" % d
                       for d in (440, 1024, 1024, 1024)]) +
        "<AffineTransform> <InputDim> 1024 <OutputDim> 1500 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04\n<Softmax> <InputDim> 1500 <OutputDim> 1500\n",
        rows=256, dim=440, K=1500, streams=None),
    "cfg2 2x Lstm(512) + Affine 512->1500, T=20 x S=100": dict(
        proto="<Lstm> <InputDim> 40 <OutputDim> 512 <ClipGradient> 5 <ParamScale> 0.01\n<Lstm> <InputDim> 512 <OutputDim> 512 <ClipGradient> 5 <ParamScale> 0.01\n"
              "<AffineTransform> <InputDim> 512 <OutputDim> 1500 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04\n<Softmax> <InputDim> 1500 <OutputDim> 1500\n",
        rows=2000, dim=40, K=1500, streams=100),
    "cfg4 FSMN 6 x (Affine 1024, ReLU, Affine 512, CompactFsmn 20/20), one 1000-frame utterance": dict(
        proto="".join(["<AffineTransform> <InputDim> %d <OutputDim> 1024 <BiasMean> 0 <BiasRange> 0.1 <ParamStddev> 0.04\n<ReLU> <InputDim> 1024 <OutputDim> 1024\n"
                       "<AffineTransform> <InputDim> 1024 <OutputDim> 512 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04\n"
                       "<CompactFsmn> <InputDim> 512 <OutputDim> 512 <PastContext> 20 <FutureContext> 20\n" % d for d in (440, 512, 512, 512, 512, 512)]) +
        "<AffineTransform> <InputDim> 512 <OutputDim> 1500 <BiasMean> 0 <BiasRange> 0 <ParamStddev> 0.04\n<Softmax> <InputDim> 1500 <OutputDim> 1500\n",
        rows=1000, dim=440, K=1500, streams=None),
}

def build_net(NN, c):
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "proto.txt")
        open(p, "w").write("<NnetProto>\n" + c["proto"] + "</NnetProto>\n")
        NN.srand(777)
        return NN.Nnet.init(p)


def run_config(name, c, reps=7):
    """One device-resident minibatch of the configuration through the handle API; returns the JSON-able result."""
    import time
    net = build_net(NN, c)
    net.set_train_options(learn_rate=1e-4, momentum=0.9)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((c["rows"], c["dim"])).astype(np.float32)
    t = rng.integers(0, c["K"], size=c["rows"]).astype(np.int32)
    if c["streams"]:
        net.reset_streams([1] * c["streams"])
    dev, _ = NN.upload(x)
    xent = NN.Xent()

    def step():
        NN.train_step_xent(net, xent, dev, t, on_device=True, rows=c["rows"], cols=c["dim"])
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        NN.device_sync(); a.record(); step(); NN.device_sync(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    NN.device_sync(); h0 = time.perf_counter()
    for _ in range(20):
        step()
    h1 = time.perf_counter(); NN.device_sync(); h2 = time.perf_counter()
    host_ms, total_ms = (h1 - h0) / 20 * 1e3, (h2 - h0) / 20 * 1e3
    res = {"config": name, "ms_per_minibatch": round(ms, 3), "frames_per_s": round(c["rows"] / ms * 1e3), "params": net.num_params,
           "host_enqueue_ms": round(host_ms, 3), "ms_back_to_back": round(total_ms, 3), "frames_per_s_back_to_back": round(c["rows"] / total_ms * 1e3),
           "step_replays": NN.step_replays()}
    net.close()
    return res


def reference_config(c, steps, warmup=1, threads=None):
    """The same minibatch through the UNMODIFIED reference classes on the host cores (oracle/_ref/ref_driver bench): frames/s."""
    import subprocess
    from oracle import kaldi_io
    root = __file__.rsplit("/tools/", 1)[0]
    drv = os.path.join(root, "oracle", "_ref", "ref_driver")
    if not os.path.exists(drv):
        return None
    cores = threads or os.cpu_count() or 1
    rng = np.random.default_rng(0)
    x = rng.standard_normal((c["rows"], c["dim"])).astype(np.float32)
    t = rng.integers(0, c["K"], size=c["rows"]).astype(np.int32)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "proto.txt"), "w").write("<NnetProto>\n" + c["proto"] + "</NnetProto>\n")
        env = dict(os.environ, OPENBLAS_NUM_THREADS=str(cores))
        subprocess.check_call([drv, "init", "proto.txt", "model.bin", "777", "1"], cwd=td, env=env, stderr=subprocess.DEVNULL)
        kaldi_io.write_mat(os.path.join(td, "input.mat"), x)
        open(os.path.join(td, "targets.txt"), "w").write(" ".join(map(str, t)) + "\n")
        spec = "input input.mat\nloss xent\ntargets targets.txt\nlearn_rate 1e-4\nmomentum 0.9\niters %d\nwarmup %d\n" % (steps + warmup, warmup)
        if c["streams"]:
            spec += "reset_flags %s\n" % ",".join(["1"] * c["streams"])
        open(os.path.join(td, "spec.txt"), "w").write(spec)
        pr = subprocess.run([drv, "bench", "model.bin", "spec.txt"], cwd=td, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    r = json.loads(pr.stdout.decode().strip().splitlines()[-1])
    return {"value": r["frames_per_sec"], "unit": "frames/s", "cores": cores, "kind": "reference",
            "sample": "%d timed minibatches of %d frames, OpenBLAS threads=%d" % (steps, c["rows"], cores)}


if __name__ == "__main__":
    NN.select_device(0)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, c in CFG.items():
        if only and not name.startswith(only):
            continue
        print(json.dumps(run_config(name, c)), flush=True)
