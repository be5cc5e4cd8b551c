"""world_size-2 test (gloo, CPU) of the data-parallel exchange step the workers implement on NCCL
(kaldi-aslp_b200/host/parallel.cc): pack every parameter tensor into ONE fp32 arena, one sum-allreduce, apply
BSP weights / the BMUF filter, plus the termination protocol (a rank that ran out of utterances keeps joining
zero-frame syncs until the GLOBAL frame count is 0).  Each rank's result is compared with the in-process N-replica
restatement of the reference formulas (oracle/aslp_oracle.py: bsp-worker.cc:33-58, bmuf-worker.cc:37-68)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import aslp_oracle as O

SHAPES = [(7, 5), (3,), (4, 12), (1,)]       # ragged tensors, as GetGpuParams hands them out
PERIODS = {0: [256, 256, 128], 1: [256]}      # rank 1 runs out of data first


def make_params(rank, step):
    rng = np.random.default_rng(1000 * rank + step)
    return [rng.standard_normal(s).astype(np.float32) for s in SHAPES]


def pack(ts):
    return np.concatenate([t.ravel() for t in ts]).astype(np.float32)


def all_finished(frames):
    """IWorker::AllFinished: int allreduce of the per-rank frame counts; true when nobody had data."""
    t = torch.tensor([frames], dtype=torch.int64)
    dist.all_reduce(t)
    return int(t.item()) == 0, int(t.item())


SEGMENTS = [(0, 2), (2, 4)]                  # tensors of "component" 0 and 1 (IWorker::InitParam(nnet) remembers these runs)


def seg_slices():
    sizes = [int(np.prod(s)) for s in SHAPES]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return [slice(int(offs[a]), int(offs[b])) for a, b in SEGMENTS]


def bmuf_apply(w, w_prev, delta_prev, g, sl):
    delta = np.float32(0.5) * delta_prev[sl] + np.float32(0.5) * np.float32(1.0) * g
    w[sl] = (w_prev[sl] + delta).astype(np.float32)
    w_prev[sl] = w[sl]
    delta_prev[sl] = delta.astype(np.float32)


def worker_pipelined(rank, world, port, out):
    """The exchange pipelined by component (IWorker::BeginSynchronize / EndSynchronize, host/parallel.cc): a rank with data sends
    its frame count first WITHOUT waiting for it, then each component's slice right behind that component's update, top
    component first, and waits for everything at the end; a rank that is out of data answers with the blocking form, cut into
    the same slices in the same order.  The collectives of the two forms must pair up and give the blocking result."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = pack(make_params(0, 0))
        w_prev, delta_prev = w.copy(), np.zeros_like(w)
        trace, step = [], 0
        my = PERIODS[rank]
        sls = seg_slices()
        while True:
            frames = my[step] if step < len(my) else 0
            if frames > 0:
                cnt = torch.tensor([frames], dtype=torch.int64)
                h_cnt = dist.all_reduce(cnt, async_op=True)            # BeginSynchronize: nobody waits for the count yet
                upd = 0.01 * pack(make_params(rank + 1, step + 1))
                pending = []
                for sl in reversed(sls):                                # Backpropagate: top component's Update first ...
                    w[sl] = w[sl] + upd[sl]
                    arena = torch.from_numpy((w[sl] - w_prev[sl]).astype(np.float32))
                    pending.append((sl, arena, dist.all_reduce(arena, async_op=True)))   # ... its exchange right behind it
                h_cnt.wait()                                            # EndSynchronize
                total = int(cnt.item())
                assert total > 0
                for sl, arena, h in pending:
                    h.wait()
                    bmuf_apply(w, w_prev, delta_prev, arena.numpy(), sl)
            else:
                done, total = all_finished(0)                           # Synchronize(0), blocking, cut the same way
                if done:
                    break
                for sl in reversed(sls):
                    arena = torch.from_numpy((w[sl] - w_prev[sl]).astype(np.float32))
                    dist.all_reduce(arena)
                    bmuf_apply(w, w_prev, delta_prev, arena.numpy(), sl)
            trace.append((frames, total, w.copy()))
            step += 1
        out[rank] = trace
    finally:
        dist.destroy_process_group()


def worker(rank, world, port, kind, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = pack(make_params(0, 0))                       # every rank starts from the same model
        w_prev, delta_prev = w.copy(), np.zeros_like(w)
        trace, step = [], 0
        my = PERIODS[rank]
        while True:
            frames = my[step] if step < len(my) else 0    # zero-frame dummy sync once out of data
            if frames > 0:                                # "local training": a rank-specific perturbation
                w = w + 0.01 * pack(make_params(rank + 1, step + 1))
            done, total = all_finished(frames)
            if done:
                break
            if kind == "bsp":
                arena = torch.from_numpy(w * np.float32(frames / total))
                dist.all_reduce(arena)
                w = arena.numpy().copy()
            else:                                          # bmuf, momentum 0.5, lr 1.0
                arena = torch.from_numpy(w - w_prev)
                dist.all_reduce(arena)
                g = arena.numpy()
                delta = np.float32(0.5) * delta_prev + np.float32(0.5) * np.float32(1.0) * g
                w = (w_prev + delta).astype(np.float32)
                w_prev, delta_prev = w.copy(), delta.astype(np.float32)
            trace.append((frames, total, w.copy()))
            step += 1
        out[rank] = trace
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def simulate(kind):
    """N-replica simulation with the oracle's formulas."""
    world = 2
    w = [pack(make_params(0, 0)) for _ in range(world)]
    w_prev, delta_prev = w[0].copy(), np.zeros_like(w[0])
    trace, step = [], 0
    while True:
        frames = [PERIODS[r][step] if step < len(PERIODS[r]) else 0 for r in range(world)]
        for r in range(world):
            if frames[r] > 0:
                w[r] = w[r] + 0.01 * pack(make_params(r + 1, step + 1))
        if sum(frames) == 0:
            break
        if kind == "bsp":
            avg = O.bsp_sync(w, frames)
            w = [avg.copy() for _ in range(world)]
        else:
            new, w_prev, delta_prev = O.bmuf_sync(w, w_prev, delta_prev, 0.5, 1.0)
            w = [new.copy() for _ in range(world)]
        trace.append((frames, w[0].copy()))
        step += 1
    return trace


@pytest.mark.parametrize("kind", ["bsp", "bmuf"])
def test_two_rank_sync_matches_replica_simulation(kind):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = free_port()
        procs = [ctx.Process(target=worker, args=(r, 2, port, kind, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0, p.exitcode
        got = {r: list(out[r]) for r in range(2)}
    want = simulate(kind)
    # both ranks took part in every sync, including rank 1's zero-frame ones (termination protocol)
    assert len(got[0]) == len(got[1]) == len(want) == 3
    for step, (frames, w_want) in enumerate(want):
        for r in range(2):
            f, total, w = got[r][step]
            assert f == frames[r] and total == sum(frames)
            np.testing.assert_allclose(w, w_want, rtol=1e-6, atol=1e-7)
    # replicas are bit-identical to each other after a sync (same allreduce result, same filter)
    for step in range(len(want)):
        assert np.array_equal(got[0][step][2], got[1][step][2])


def test_two_rank_pipelined_exchange_pairs_with_the_blocking_form():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = free_port()
        procs = [ctx.Process(target=worker_pipelined, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0, p.exitcode
        got = {r: list(out[r]) for r in range(2)}
    want = simulate("bmuf")
    assert len(got[0]) == len(got[1]) == len(want) == 3
    for step, (frames, w_want) in enumerate(want):
        for r in range(2):
            f, total, w = got[r][step]
            assert f == frames[r] and total == sum(frames)
            np.testing.assert_allclose(w, w_want, rtol=1e-6, atol=1e-7)
        assert np.array_equal(got[0][step][2], got[1][step][2])
