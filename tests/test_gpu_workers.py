"""IWorker parity (BSP / BMUF / SOD over NCCL, kaldi-aslp_b200/host/parallel.cc) against the N-replica restatement of
src/aslp-parallel/{bsp,bmuf,sod}-worker.cc in oracle/aslp_oracle.py (the reference has no tests for these and needs MPI; the
restatement itself is pinned against the reference's own worker classes run over a stand-in for mpi.h,
tests/test_cpu_oracle_pinning.py::test_worker_restatements_match_the_reference_workers).  One rank on any GPU box; two ranks when the box has two GPUs."""
import os
import tempfile
import time

import numpy as np
import pytest

from oracle import aslp_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dnn_xent")
FRAMES = {0: 96, 1: 32}


def local_step(NN, net, rank):
    """one frame-CE minibatch on rank-specific data so that the replicas diverge"""
    rng = np.random.default_rng(100 + rank)
    x = rng.standard_normal((24, 20)).astype(np.float32)
    t = rng.integers(0, 16, size=24).astype(np.int32)
    net.set_train_options(0.05, 0.0, 0.0, 0.0)
    NN.train_step_xent(net, NN.Xent(), x, t)


def rank_main(rank, world, kind, id_file, out_dir, mode="blocking"):
    import kaldi_aslp_b200 as K
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
    w0 = net.get_params()
    worker = NN.Worker(kind, ident, world, rank, bmuf_momentum=0.5, bmuf_learn_rate=1.0, sod_solver="momentum")
    # blocking: the reference's form.  segmented: tensors registered by component, every exchange goes component by component.
    # overlapped: begin_synchronize before the minibatch, each component exchanged behind its Update, end_synchronize after.
    if mode == "blocking":
        worker.init_param(net)
    else:
        worker.init_param_by_component(net)
    if mode == "overlapped":
        assert worker.can_overlap()
    trace = [w0]
    for step in range(2):
        # one rank falls back to the blocking call in the second round: the collectives of the two forms must pair up
        if mode == "overlapped" and not (step == 1 and rank == world - 1 and world > 1):
            worker.begin_synchronize(FRAMES[rank])
            local_step(NN, net, rank + 10 * step)
            assert worker.end_synchronize() is True
            trace.append(net.get_params())             # the model before the exchange is never visible here: run_world compares
            trace.append(net.get_params())             # the result with the blocking run's
            continue
        local_step(NN, net, rank + 10 * step)
        trace.append(net.get_params())                 # before the sync
        assert worker.synchronize(FRAMES[rank]) is True
        trace.append(net.get_params())                 # after
    assert worker.synchronize(0) is False              # every rank out of data -> protocol ends
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.stack(trace))
    worker.close()
    net.close()


def expected(kind, traces):
    """replay with the restated formulas: traces[r] = [w0, pre1, post1, pre2, post2]; returns want[step][rank]"""
    world = len(traces)
    w_prev = traces[0][0].copy()
    delta_prev = np.zeros_like(w_prev)
    prev_r = [traces[r][0].copy() for r in range(world)]      # SOD keeps a per-rank previous model
    s1 = np.zeros_like(w_prev); s2 = np.zeros_like(w_prev)
    out = []
    for step in range(2):
        pre = [traces[r][1 + 2 * step] for r in range(world)]
        if kind == "bsp":
            w = O.bsp_sync(pre, [FRAMES[r] for r in range(world)])
            out.append([w] * world)
        elif kind == "bmuf":
            w, w_prev, delta_prev = O.bmuf_sync(pre, w_prev, delta_prev, 0.5, 1.0)
            out.append([w] * world)
        else:
            # sod-worker.cc:46-60: G = SUM_r (prev_r - w_r); the optimizer moves each rank's OWN current weights by the
            # common update (the replicas are NOT averaged), then prev_r = new w_r
            g = sum(prev_r[r] - pre[r] for r in range(world)).astype(np.float32)
            new = []
            s1n = s2n = None
            for r in range(world):
                w, s1n, s2n = O.sod_optimize("momentum", pre[r], g, s1, s2, 0.01, 0.9, 0.0, step + 1)
                new.append(w)
            s1, s2 = s1n, s2n
            prev_r = [w.copy() for w in new]
            out.append(new)
    return out


def run_world(world, kind, mode="blocking"):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        id_file = os.path.join(d, "nccl_id")
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=rank_main, args=(r, world, kind, id_file, d, mode)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0, (kind, mode, p.exitcode)
        traces = [np.load(os.path.join(d, "rank%d.npy" % r)) for r in range(world)]
    if mode == "overlapped":
        # steps are deterministic and a two-term sum does not depend on how the all-reduce is cut up: bit-identical to blocking
        ref = run_world(world, kind, "blocking")
        for step in range(2):
            for r in range(world):
                assert np.array_equal(traces[r][2 + 2 * step], ref[r][2 + 2 * step]), (kind, step, r)
        return traces
    want = expected(kind, traces)
    for step in range(2):
        for r in range(world):
            got = traces[r][2 + 2 * step]
            err = np.max(np.abs(got - want[step][r])) / np.max(np.abs(want[step][r]))
            assert err < 1e-6, (kind, step, r, err)
        if world > 1 and kind != "sod":      # replicas agree bit for bit after a BSP / BMUF sync
            assert np.array_equal(traces[0][2 + 2 * step], traces[1][2 + 2 * step]), (kind, step)
    return traces


@pytest.mark.parametrize("kind", ["bsp", "bmuf", "sod"])
def test_single_rank_worker_matches_restated_formulas(kind):
    run_world(1, kind)


@pytest.mark.parametrize("kind", ["bsp", "bmuf", "sod"])
def test_two_rank_worker_matches_replica_simulation(kind):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, kind)


@pytest.mark.parametrize("kind,mode", [("bsp", "segmented"), ("bsp", "overlapped"), ("bmuf", "segmented"), ("bmuf", "overlapped"), ("sod", "overlapped")])
def test_single_rank_exchange_by_component(kind, mode):
    """IWorker::InitParam(nnet) / BeginSynchronize / EndSynchronize: the exchange pipelined by layer gives what the blocking one does"""
    run_world(1, kind, mode)


@pytest.mark.parametrize("kind,mode", [("bsp", "segmented"), ("bsp", "overlapped"), ("bmuf", "overlapped"), ("sod", "overlapped")])
def test_two_rank_exchange_by_component(kind, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    run_world(2, kind, mode)
