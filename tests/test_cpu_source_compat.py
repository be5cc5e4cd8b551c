"""SURVEY 8(b), first row: the C++ component / Nnet / loss / worker interface is SOURCE-compatible with the reference's.
The UNMODIFIED reference mains (read where they lie under /root/reference/src, never copied) must compile against
kaldi-aslp_b200/host through the include-path shims of kaldi-aslp_b200/compat.  Needs the reference tree, so it runs in the
build container only (skipped on the GPU box); tests/test_gpu_dropin_mains.py then RUNS the binaries built from the same
sources by `make -C oracle dropin`."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"

MAINS = [
    "aslp-nnetbin/aslp-nnet-train-warp-ctc-streams.cc",
    "aslp-nnetbin/aslp-nnet-train-blstm-streams-lc.cc",
    "aslp-nnetbin/aslp-nnet-train-frame.cc",
    "aslp-nnetbin/aslp-nnet-train-lstm-streams.cc",
    "aslp-nnetbin/aslp-nnet-train-ctc-streams.cc",
    "aslp-nnetbin/aslp-nnet-train-blstm-parallel.cc",
    "aslp-nnetbin/aslp-nnet-info.cc",
    "aslp-nnetbin/aslp-nnet-copy.cc",
    "aslp-nnetbin/aslp-nnet-init.cc",
    "aslp-nnetbin/aslp-nnet-train-perutt.cc",
    "aslp-nnetbin/aslp-nnet-train-mse.cc",
    "aslp-nnetbin/aslp-nnet-train-blstm-streams.cc",
    "aslp-nnetbin/aslp-nnet-train-lstm-streams-skip.cc",
    "aslp-nnetbin/aslp-nnet-forward.cc",
    "aslp-nnetbin/aslp-nnet-forward-skip.cc",
    "aslp-nnetbin/aslp-nnet-forward-blstm-lc.cc",
    "aslp-nnetbin/aslp-nnet-forward-mimo.cc",
    "aslp-nnetbin/aslp-nnet-train-frame-mimo.cc",
    "aslp-parallelbin/aslp-nnet-train-lc-blstm-streams-worker.cc",
    "aslp-parallelbin/aslp-nnet-train-frame-worker.cc",
    "aslp-parallelbin/aslp-nnet-train-lstm-stream-worker.cc",
]


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("main", MAINS)
def test_unmodified_reference_main_compiles_against_host_layer(main):
    pkg = os.path.join(ROOT, "kaldi-aslp_b200")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-I", os.path.join(pkg, "compat"), "-I", os.path.join(pkg, "host"),
           "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", os.path.join(REF_SRC, main)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]


def test_compat_shims_only_forward():
    """the shim directory holds include forwards only: no reference text, no code"""
    compat = os.path.join(ROOT, "kaldi-aslp_b200", "compat")
    n = 0
    for dp, _, fns in os.walk(compat):
        for fn in fns:
            lines = [l for l in open(os.path.join(dp, fn)).read().splitlines() if l.strip() and not l.startswith("//")]
            assert len(lines) == 1 and lines[0].startswith('#include "../../host/'), (fn, lines)
            n += 1
    assert n >= 20
