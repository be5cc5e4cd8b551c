"""Async parameter-server modes (EASGD / ASGD / MASGD, kaldi-aslp_b200/host/parallel-async.cc) over NCCL send / recv: rank 0
serves, rank 1 trains.  With one worker the arrival order is fixed, so the run replays exactly against the restated formulas
of oracle/aslp_oracle.py (easgd-*.cc, asgd-*.cc, masgd-server.cc; the reference has no tests for these and needs MPI -- the
restatement itself is pinned against the reference's own server and worker classes run over a stand-in for mpi.h,
tests/test_cpu_oracle_pinning.py::test_async_restatements_match_the_reference_servers).  Needs two GPUs (NCCL does not put two ranks on one device)."""
import os
import tempfile
import time

import numpy as np
import pytest

from oracle import aslp_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dnn_xent")
NSYNC = 3


def local_step(NN, net, seed):
    rng = np.random.default_rng(500 + seed)
    x = rng.standard_normal((24, 20)).astype(np.float32)
    t = rng.integers(0, 16, size=24).astype(np.int32)
    net.set_train_options(0.05, 0.0, 0.0, 0.0)
    NN.train_step_xent(net, NN.Xent(), x, t)


def rank_main(rank, kind, id_file, out_dir, port):
    os.environ["ASLP_CTRL_PORT"] = str(port)
    from kaldi_aslp_b200 import nnet as NN
    NN.select_device(rank)
    if rank == 0:
        ident = NN.nccl_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(ident)
        os.replace(id_file + ".tmp", id_file)
    else:
        for _ in range(600):
            if os.path.exists(id_file):
                break
            time.sleep(0.1)
        ident = open(id_file, "rb").read()
    net = NN.Nnet.read(os.path.join(GOLD, "model.bin"))
    if rank == 0:
        # the server starts from a DIFFERENT model than the worker so that every term of the formulas is exercised
        local_step(NN, net, 99)
        w_server0 = net.get_params()
        server = NN.Server(kind, ident, 2, alpha=0.3, sync_period=2, momentum=0.5)
        server.init_param(net)
        server.run()                                   # returns when the worker has sent kMsgFinished
        np.save(os.path.join(out_dir, "server.npy"), np.stack([w_server0, net.get_params()]))
        server.close()
    else:
        worker = NN.Worker(kind, ident, 2, 1, bmuf_learn_rate=0.3)      # alpha of the EASGD worker
        worker.init_param(net)
        trace = [net.get_params()]
        for step in range(NSYNC):
            local_step(NN, net, step)
            trace.append(net.get_params())             # before the sync
            assert worker.synchronize(24) is True
            trace.append(net.get_params())             # after
        worker.stop()
        np.save(os.path.join(out_dir, "worker.npy"), np.stack(trace))
        worker.close()
    net.close()


@pytest.mark.parametrize("kind", ["easgd", "asgd", "masgd"])
def test_one_worker_one_server_matches_restated_formulas(kind):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with tempfile.TemporaryDirectory() as d:
        id_file = os.path.join(d, "nccl_id")
        port = 29700 + (os.getpid() % 200) + {"easgd": 0, "asgd": 1, "masgd": 2}[kind] * 211
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=rank_main, args=(r, kind, id_file, d, port)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(240)
            if p.is_alive():
                p.kill()
            assert p.exitcode == 0, (kind, p.exitcode)
        server = np.load(os.path.join(d, "server.npy"))
        trace = np.load(os.path.join(d, "worker.npy"))
    w_server = server[0].copy()
    w_prev = trace[0].copy()
    diff = np.zeros_like(w_server)
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    for step in range(NSYNC):
        pre, post = trace[1 + 2 * step], trace[2 + 2 * step]
        if kind == "easgd":
            want_worker, w_server = O.easgd_exchange(pre, w_server, 0.3)
        elif kind == "asgd":
            # one worker and sync_period = 2: from the second update on the worker is "waited" and released at once (it is the
            # only running worker), so it always receives the server's new model
            want_worker, w_server = O.asgd_update(pre, w_prev, w_server, 0.3)
        else:
            want_worker, w_server, diff = O.masgd_update(pre, w_prev, w_server, diff, 0.5)
        w_prev = want_worker.copy()
        assert rel(post, want_worker) < 1e-6, (kind, step, rel(post, want_worker))
    assert rel(server[1], w_server) < 1e-6, (kind, rel(server[1], w_server))
    assert np.max(np.abs(server[1] - server[0])) > 1e-4 * np.max(np.abs(server[0]))       # the server really moved
