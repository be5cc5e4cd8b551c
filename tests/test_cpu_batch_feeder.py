"""Host logic of the minibatch feeder (kaldi-aslp_b200/host/batch-feeder.h) without a device: a small C++
driver is compiled against the header and run at several ASLP_FEEDER_DEPTH values -- in-order delivery with slow producers
and consumers, slot reuse, the end-of-data signal, Join() publishing what the feeder thread wrote, and an exception thrown
by fill() on the feeder thread surfacing in Next().  (The device side -- page-locked slots recycled behind a CUDA event --
is covered by the depth-invariance tests in tests/test_gpu_cli.py.)"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kaldi-aslp_b200")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    if not os.path.exists(os.path.join(PKG, "libaslp_nnet.so")):
        pytest.skip("libaslp_nnet.so is not built (run __graft_entry__.build())")
    exe = str(tmp_path_factory.mktemp("feeder") / "batch_feeder_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "batch_feeder_test.cc"), "-o", exe,
                           "-L", PKG, "-laslp_nnet", "-laslp_b200", "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"])
    return exe


@pytest.mark.parametrize("depth", ["0", "1", "2", "3", "8"])
def test_batches_arrive_in_order_at_every_depth(driver, depth):
    r = subprocess.run([driver, "200"], env=dict(os.environ, ASLP_FEEDER_DEPTH=depth), stdout=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert r.stdout.strip() == "OK depth=%s delivered=200 sum=%d" % (depth, 199 * 200 // 2)


@pytest.mark.parametrize("depth", ["0", "2"])
def test_empty_table(driver, depth):
    r = subprocess.run([driver, "0"], env=dict(os.environ, ASLP_FEEDER_DEPTH=depth), stdout=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and "delivered=0" in r.stdout, r.stdout


@pytest.mark.parametrize("depth", ["0", "2"])
def test_exception_on_the_feeder_thread_reaches_the_consumer(driver, depth):
    r = subprocess.run([driver, "50", "17"], env=dict(os.environ, ASLP_FEEDER_DEPTH=depth), stdout=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 3 and "fill failed on purpose" in r.stdout, r.stdout


def test_randomizer_replay_predicts_the_reference_refill_grouping(tmp_path):
    """FrameDataReader prepares the next randomizer refill ahead of time from frame counts alone (RandomizerReplay); 2000 random
    tables, including ones that end exactly on a refill boundary, against the loop that asks the real randomizer."""
    if not os.path.exists(os.path.join(PKG, "libaslp_nnet.so")):
        pytest.skip("libaslp_nnet.so is not built (run __graft_entry__.build())")
    exe = str(tmp_path / "randomizer_replay_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "randomizer_replay_test.cc"), "-o", exe,
                           "-L", PKG, "-laslp_nnet", "-laslp_b200", "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"])
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("OK trials=2000") and "boundary_cases=0" not in r.stdout, r.stdout
