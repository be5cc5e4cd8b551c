"""Table I/O of the trainer mains without a device (kaldi-aslp_b200/host/table.h): Kaldi's extended file names -- paths, "-",
"command |" on the reading side and "| command" on the writing side, as every reference recipe passes its features
(aslp_scripts/aslp_nnet/prepare_feats_ali.sh: feats_tr="ark:copy-feats scp:... ark:- | apply-cmvn ... |") -- and compressed
feature matrices (CM / CM2).  The compressed fixtures under tests/golden/io were written by the REFERENCE's CompressedMatrix
(oracle/ref_driver.cc `compress`), together with what its own CopyToMat decodes them to."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import kaldi_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kaldi-aslp_b200")
GOLD = os.path.join(ROOT, "tests", "golden", "io")


@pytest.fixture(scope="module")
def copy_tool(tmp_path_factory):
    if not os.path.exists(os.path.join(PKG, "libaslp_nnet.so")):
        pytest.skip("libaslp_nnet.so is not built (run __graft_entry__.build())")
    exe = str(tmp_path_factory.mktemp("tableio") / "table_io_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "cpp", "table_io_test.cc"), "-o", exe,
                           "-L", PKG, "-laslp_nnet", "-laslp_b200", "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64", "-lcudart",
                           "-Wl,-rpath,/usr/local/cuda/lib64", "-lpthread"])
    return exe


def write_ark(path, utts):
    with open(path, "wb") as f:
        for key, m in utts:
            m = np.ascontiguousarray(m, np.float32)
            f.write(key.encode() + b" \0BFM " + b"\x04" + struct.pack("<i", m.shape[0]) + b"\x04" + struct.pack("<i", m.shape[1]) + m.tobytes())


def read_ark(path):
    out, data, pos = {}, open(path, "rb").read(), 0
    while pos < len(data):
        sp = data.index(b" ", pos)
        key = data[pos:sp].decode()
        assert data[sp + 1:sp + 6] == b"\0BFM ", data[sp + 1:sp + 8]
        p = sp + 6
        rows, cols = struct.unpack("<i", data[p + 1:p + 5])[0], struct.unpack("<i", data[p + 6:p + 10])[0]
        p += 10
        out[key] = np.frombuffer(data, np.float32, rows * cols, p).reshape(rows, cols).copy()
        pos = p + rows * cols * 4
    return out


@pytest.fixture()
def table(tmp_path):
    rng = np.random.default_rng(1)
    utts = [("utt%d" % i, rng.standard_normal((3 + i, 4)).astype(np.float32)) for i in range(5)]
    path = str(tmp_path / "in.ark")
    write_ark(path, utts)
    return path, dict(utts)


def same(got, want):
    assert list(got) == list(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


def test_file_to_file(copy_tool, table, tmp_path):
    path, want = table
    out = str(tmp_path / "out.ark")
    subprocess.check_call([copy_tool, "ark:" + path, "ark:" + out])
    same(read_ark(out), want)


def test_piped_rspecifier_as_the_recipes_pass_it(copy_tool, table, tmp_path):
    path, want = table
    out = str(tmp_path / "out.ark")
    subprocess.check_call([copy_tool, "ark:cat %s |" % path, "ark:" + out])
    same(read_ark(out), want)
    # a two-stage pipe with option letters, trailing blanks after the bar
    subprocess.check_call([copy_tool, "ark,s,cs:cat %s | cat - | " % path, "ark:" + out])
    same(read_ark(out), want)


def test_stdin_and_stdout(copy_tool, table, tmp_path):
    path, want = table
    out = str(tmp_path / "out.ark")
    with open(path, "rb") as fin, open(out, "wb") as fout:
        subprocess.check_call([copy_tool, "ark:-", "ark:-"], stdin=fin, stdout=fout)
    same(read_ark(out), want)


def test_piped_wspecifier(copy_tool, table, tmp_path):
    path, want = table
    out = str(tmp_path / "out.ark")
    subprocess.check_call([copy_tool, "ark:" + path, "ark:| cat > %s" % out])
    same(read_ark(out), want)


def test_scp_with_offsets_and_piped_entries(copy_tool, table, tmp_path):
    path, want = table
    data = open(path, "rb").read()
    scp = str(tmp_path / "in.scp")
    with open(scp, "w") as f:
        for i, k in enumerate(want):
            off = data.index(k.encode() + b" ") + len(k) + 1
            if i % 2 == 0:
                f.write("%s %s:%d\n" % (k, path, off))
            else:                                     # "key command |": the object comes from a command's output
                f.write("%s tail -c +%d %s |\n" % (k, off + 1, path))
    out = str(tmp_path / "out.ark")
    subprocess.check_call([copy_tool, "scp:" + scp, "ark:" + out])
    same(read_ark(out), want)


@pytest.mark.parametrize("name,key", [("cm1", "uttA"), ("cm2", "uttB")])
def test_compressed_matrices_decode_like_the_reference(copy_tool, name, key, tmp_path):
    out = str(tmp_path / "out.ark")
    subprocess.check_call([copy_tool, "ark:" + os.path.join(GOLD, name + ".ark"), "ark:" + out])
    got = read_ark(out)
    want = kaldi_io.read(os.path.join(GOLD, name + ".ark.decoded"))
    assert list(got) == [key] and got[key].shape == want.shape
    assert np.array_equal(got[key], want)            # same formulas in the same float arithmetic: bit-exact


def test_missing_file_and_failing_command_are_loud(copy_tool, tmp_path):
    r = subprocess.run([copy_tool, "ark:/nonexistent/feats.ark", "ark:" + str(tmp_path / "o.ark")], stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Cannot open" in r.stderr
