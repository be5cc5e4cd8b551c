"""Build recipe for the native parts (no setup.py, no JIT cache: everything lands in-tree).

  libaslp_b200.so  -- csrc/*.cu, hand-written sm_100a kernels behind include/aslp_b200.h + include/ctc.h
  libaslp_nnet.so  -- host/*.cc, the C++ Component/Nnet/IWorker mirror + its C handle API (include/aslp_nnet_c.h)
  bin/*            -- trainer CLIs with the reference's flags

nvcc cross-compiles for sm_100a without a GPU.  Objects are cached by source mtime.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
BIN_SRC = os.path.join(HERE, "bin")
OBJ = os.path.join(HERE, "build")
LIB_CUDA = os.path.join(HERE, "libaslp_b200.so")
LIB_HOST = os.path.join(HERE, "libaslp_nnet.so")
BIN_OUT = os.path.join(HERE, "build", "bin")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-sign-compare", "-I", os.path.join(ROOT, "include"), "-I", HOST,
             "-I", "/usr/local/cuda/include"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src, *extra))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_cuda(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if _newer(s, o, hdrs):
            jobs.append([NVCC, *NVCC_FLAGS, "-c", s, "-o", o])
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out.strip():
                print(out)
    if jobs or not os.path.exists(LIB_CUDA):
        _run([NVCC, "-shared", "-o", LIB_CUDA, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
              "-lnccl", "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
    return LIB_CUDA


def build_host(verbose=False):
    srcs = sorted(glob.glob(os.path.join(HOST, "*.cc")))
    if not srcs:
        return None
    os.makedirs(OBJ, exist_ok=True)
    hdrs = glob.glob(os.path.join(HOST, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, "host_" + os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if _newer(s, o, hdrs):
            jobs.append(["g++", *CXX_FLAGS, "-c", s, "-o", o])
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out.strip():
                print(out)
    if jobs or not os.path.exists(LIB_HOST):
        _run(["g++", "-shared", "-o", LIB_HOST, *objs, "-L", HERE, "-laslp_b200", "-Wl,-rpath,$ORIGIN",
              "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"])
    # CLI trainers
    os.makedirs(BIN_OUT, exist_ok=True)
    bjobs = []
    for s in sorted(glob.glob(os.path.join(BIN_SRC, "*.cc"))):
        exe = os.path.join(BIN_OUT, os.path.basename(s)[:-3])
        if _newer(s, exe, hdrs + glob.glob(os.path.join(BIN_SRC, "*.h")) + [LIB_HOST]):
            bjobs.append(["g++", *CXX_FLAGS, "-I", BIN_SRC, s, "-o", exe, "-L", HERE, "-laslp_nnet", "-laslp_b200",
                          "-Wl,-rpath,$ORIGIN/../..", "-L/usr/local/cuda/lib64", "-lcudart",
                          "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"])
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, bjobs))
    return LIB_HOST


def build_all(verbose=False):
    build_cuda(verbose)
    build_host(verbose)


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv)
    print("built", LIB_CUDA, LIB_HOST if os.path.exists(LIB_HOST) else "")
