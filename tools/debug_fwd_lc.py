"""Debug aid: per-row difference of aslp-nnet-forward-blstm-lc against the reference archive."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_cli import read_feats_ark, BIN, GOLD
d = os.path.join(GOLD, "cli_fwd")
np.set_printoptions(linewidth=250, precision=2)
for fold in ("1", "0"):
    env = dict(os.environ, ASLP_LSTM_FOLD_PROJECTION=fold)
    exe, case, flags = "aslp-nnet-forward-blstm-lc", "cli_lc", ["--chunk-size=8", "--right-splice=3", "--apply-log=false"]
    out = "/tmp/dbg_fold%s.ark" % fold
    r = subprocess.run([os.path.join(BIN, exe)] + flags + [os.path.join(GOLD, case, "ref_out.nnet"), "ark:" + os.path.join(GOLD, case, "feats.ark"), "ark:" + out],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    print("fold", fold, "rc", r.returncode, r.stdout[-200:] if r.returncode else "")
    want, got = read_feats_ark(os.path.join(d, "lc_chunks.ark")), read_feats_ark(out)
    for k in list(want)[:3]:
        w, g = np.exp(want[k]), got[k]
        print(k, w.shape, "per-row max|d| x1e6:", (np.abs(g - w).max(axis=1) * 1e6).round(1))
