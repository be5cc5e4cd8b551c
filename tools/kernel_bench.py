"""Per-kernel-class HBM measurement of the memory-bound kernels on the aslp-nnet path:
achieved ALGORITHMIC GB/s against the measured HBM peak of MEASURED_PEAKS.json.

Each kernel is called through the C-ABI of libaslp_b200.so at a BASELINE-config geometry.  Operands rotate through
enough distinct buffer sets that the bytes touched between two uses of a set exceed twice the 126 MB L2, so every launch
streams from HBM; time = CUDA events around `reps` back-to-back launches on the launching stream, after warm-up.

  python tools/kernel_bench.py            # table + JSON lines (profiles/rNN_hbm_kernels.jsonl is a copy of the output)
  python tools/kernel_bench.py --once K   # one launch per kernel whose name contains K (for ncu captures)
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import kaldi_aslp_b200 as K  # noqa: E402

P = ctypes.c_void_p
L = K.cuda_lib()
ROOT = __file__.rsplit("/tools/", 1)[0]
L2_BYTES = 126e6


def peak_gbs():
    try:
        m = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_gbps", "hbm_GBps"):
            if k in m:
                return float(m[k])
        for v in m.values():
            if isinstance(v, dict):
                for k2, v2 in v.items():
                    if "hbm" in k2.lower() and isinstance(v2, (int, float)):
                        return float(v2)
    except Exception:  # noqa: BLE001
        pass
    return 6457.4


def stream():
    return P(torch.cuda.current_stream().cuda_stream)


def p(t):
    return P(t.data_ptr())


class Case:
    """name, algorithmic bytes per launch, bytes one buffer set occupies, make(set index) -> closure launching the kernel"""

    def __init__(self, name, geom, alg_bytes, set_bytes, make):
        self.name, self.geom, self.alg_bytes, self.set_bytes, self.make = name, geom, alg_bytes, set_bytes, make


def cases():
    cs = []
    f32 = dict(device="cuda", dtype=torch.float32)

    # ---- activations: cfg1 hidden layer class, one randomizer block of frames (32768 x 1024)
    R, D = 32768, 1024
    for kind, nm in ((0, "sigmoid"), (1, "tanh"), (2, "relu")):
        def mk_f(i, kind=kind):
            x = torch.randn(R, D, **f32); y = torch.empty_like(x)
            return lambda: K.check(L.aslp_act_fwd(stream(), kind, p(y), D, p(x), D, R, D)), (x, y)
        cs.append(Case("act_fwd_" + nm, "%dx%d" % (R, D), 8 * R * D, 8 * R * D, mk_f))

        def mk_b(i, kind=kind):
            y = torch.rand(R, D, **f32); e = torch.randn(R, D, **f32); d = torch.empty_like(y)
            return lambda: K.check(L.aslp_act_bwd(stream(), kind, p(d), D, p(y), D, p(e), D, R, D)), (y, e, d)
        cs.append(Case("act_bwd_" + nm, "%dx%d" % (R, D), 12 * R * D, 12 * R * D, mk_b))

    # ---- softmax: cfg1 output layer (K = 1500), cfg3 output (K = 72)
    for Rs, Ks in ((32768, 1500), (262144, 72)):
        def mk(i, Rs=Rs, Ks=Ks):
            x = torch.randn(Rs, Ks, **f32); y = torch.empty_like(x)
            return lambda: K.check(L.aslp_softmax_rows(stream(), p(y), Ks, p(x), Ks, Rs, Ks)), (x, y)
        cs.append(Case("softmax_rows", "%dx%d" % (Rs, Ks), 8 * Rs * Ks, 8 * Rs * Ks, mk))

    # ---- xent (sparse targets): read y, write diff
    Rs, Ks = 32768, 1500
    def mk_x(i, Rs=Rs, Ks=Ks):
        y = torch.softmax(torch.randn(Rs, Ks, **f32), 1); d = torch.empty_like(y)
        ti = torch.randint(0, Ks, (Rs,), device="cuda", dtype=torch.int32)
        tw = torch.ones(Rs, **f32); fw = torch.ones(Rs, **f32)
        st = torch.zeros(5, device="cuda", dtype=torch.float64)
        return lambda: K.check(L.aslp_xent_sparse(stream(), p(d), Ks, p(y), Ks, Rs, Ks, p(ti), p(tw), p(fw), p(st))), (y, d, ti, tw, fw, st)
    cs.append(Case("xent_sparse", "%dx%d" % (Rs, Ks), 8 * Rs * Ks, 8 * Rs * Ks, mk_x))

    # ---- axpby (AddMat with beta: read 2, write 1), col_sum (bias gradient: read 1)
    def mk_a(i):
        a = torch.randn(R, D, **f32); b = torch.randn(R, D, **f32)
        return lambda: K.check(L.aslp_axpby(stream(), p(a), D, p(b), D, R, D, 0.5, 0.9)), (a, b)
    cs.append(Case("axpby", "%dx%d" % (R, D), 12 * R * D, 8 * R * D, mk_a))

    def mk_cs(i):
        a = torch.randn(R, D, **f32); v = torch.zeros(D, **f32)
        return lambda: K.check(L.aslp_col_sum(stream(), p(v), p(a), D, R, D, 1.0, 0.9, 0.0)), (a, v)
    cs.append(Case("col_sum", "%dx%d" % (R, D), 4 * R * D, 4 * R * D, mk_cs))

    def mk_cd(i):
        a = torch.randn(R, D, **f32); b = torch.randn(R, D, **f32); v = torch.zeros(D, **f32)
        return lambda: K.check(L.aslp_col_dot(stream(), p(v), p(a), D, p(b), D, R, D, 1.0, 0.9, 0.0)), (a, b, v)
    cs.append(Case("col_dot", "%dx%d" % (R, D), 8 * R * D, 8 * R * D, mk_cd))

    # ---- BatchNorm (not on a BASELINE config).  Algorithmic bytes per SURVEY 8(d): fwd 8 B/elem, bwd 16 B/elem; both need the
    # column statistics of the WHOLE minibatch before the elementwise pass, so a matrix larger than L2 is read twice
    # (12 and 20 B/elem moved): the ceiling of `frac` at this size is 0.67 / 0.80
    def mk_bn(i):
        x = torch.randn(R, D, **f32); o = torch.empty_like(x)
        sc = torch.ones(D, **f32); sh = torch.zeros(D, **f32); mean = torch.zeros(D, **f32); istd = torch.zeros(D, **f32)
        am = torch.zeros(D, device="cuda", dtype=torch.float64); av = torch.zeros(D, device="cuda", dtype=torch.float64)
        return (lambda: K.check(L.aslp_bn_fwd_train(stream(), p(o), D, None, 0, p(x), D, R, D, p(sc), p(sh), 1e-7, p(mean), p(istd), p(am), p(av))),
                (x, o, sc, sh, mean, istd, am, av))
    cs.append(Case("bn_fwd_train", "%dx%d" % (R, D), 8 * R * D, 8 * R * D, mk_bn))

    def mk_bnb(i):
        x = torch.randn(R, D, **f32); dy = torch.randn(R, D, **f32); dx = torch.empty_like(x)
        sc = torch.ones(D, **f32); mean = torch.zeros(D, **f32); istd = torch.ones(D, **f32)
        ds = torch.zeros(D, **f32); dsh = torch.zeros(D, **f32)
        return (lambda: K.check(L.aslp_bn_bwd(stream(), p(dx), D, p(x), D, None, 0, p(dy), D, R, D, p(sc), p(mean), p(istd), 0.9, p(ds), p(dsh))),
                (x, dy, dx, sc, mean, istd, ds, dsh))
    cs.append(Case("bn_bwd", "%dx%d" % (R, D), 16 * R * D, 12 * R * D, mk_bnb))

    # ---- CNN front end of the CTC recipes (run_ctc_cnn_1dnn_2blstm.sh:67-68): 40 x 11 spliced input, 9-wide patches -> 32 positions x 128
    # filters, pooled 4:1.  Algorithmic bytes: gather reads the input once and writes the patches; scatter reads the patch
    # derivatives once and writes the input derivative; pooling reads / writes every operand once.
    Rc, NSc, STc, PDc, NPc, NFc = 16384, 11, 40, 9, 32, 128
    FDc = NSc * PDc
    LDP = (FDc + 3) // 4 * 4
    def mk_cg(i):
        x = torch.randn(Rc, NSc * STc, **f32); pt = torch.empty(Rc * NPc, LDP, **f32)
        return lambda: K.check(L.aslp_conv_gather_patches(stream(), p(pt), LDP, p(x), NSc * STc, Rc, NPc, NSc, PDc, 1, STc)), (x, pt)
    cs.append(Case("conv_gather", "%dx%d->%dx%d" % (Rc, NSc * STc, Rc * NPc, FDc), 4 * Rc * (NSc * STc + NPc * FDc), 4 * Rc * (NSc * STc + NPc * LDP), mk_cg))

    def mk_csc(i):
        pt = torch.randn(Rc * NPc, LDP, **f32); d = torch.empty(Rc, NSc * STc, **f32)
        return lambda: K.check(L.aslp_conv_scatter_patch_diffs(stream(), p(d), NSc * STc, p(pt), LDP, Rc, NPc, NSc, PDc, 1, STc)), (pt, d)
    cs.append(Case("conv_scatter", "%dx%d->%dx%d" % (Rc * NPc, FDc, Rc, NSc * STc), 4 * Rc * (NSc * STc + NPc * FDc), 4 * Rc * (NSc * STc + NPc * LDP), mk_csc))

    PSZ, NPO = 4, 8                                  # 32 positions -> 8 pools of 4, stride = 128 filters
    def mk_mp(i):
        x = torch.randn(Rc, NPc * NFc, **f32); o = torch.empty(Rc, NPO * NFc, **f32)
        return lambda: K.check(L.aslp_maxpool_fwd(stream(), p(o), NPO * NFc, p(x), NPc * NFc, Rc, NPO, PSZ, PSZ, NFc)), (x, o)
    cs.append(Case("maxpool_fwd", "%dx%d->%d" % (Rc, NPc * NFc, NPO * NFc), 4 * Rc * NFc * (NPc + NPO), 4 * Rc * NFc * (NPc + NPO), mk_mp))

    def mk_mpb(i):
        x = torch.randn(Rc, NPc * NFc, **f32); o = torch.empty(Rc, NPO * NFc, **f32); od = torch.randn(Rc, NPO * NFc, **f32)
        d = torch.empty(Rc, NPc * NFc, **f32)
        K.check(L.aslp_maxpool_fwd(stream(), p(o), NPO * NFc, p(x), NPc * NFc, Rc, NPO, PSZ, PSZ, NFc))
        return (lambda: K.check(L.aslp_maxpool_bwd(stream(), p(d), NPc * NFc, p(x), NPc * NFc, p(o), NPO * NFc, p(od), NPO * NFc, Rc, NPc, NPO, PSZ, PSZ, NFc)),
                (x, o, od, d))
    cs.append(Case("maxpool_bwd", "%dx%d" % (Rc, NPc * NFc), 4 * Rc * NFc * (2 * NPc + 2 * NPO), 4 * Rc * NFc * (2 * NPc + 2 * NPO), mk_mpb))

    # ---- Splice: 40-dim, offsets -5..5 (cfg1 front end), one randomizer block of frames
    Rs, Ds, NO = 262144, 40, 11
    def mk_sp(i):
        x = torch.randn(Rs, Ds, **f32); o = torch.empty(Rs, Ds * NO, **f32)
        off = torch.arange(-5, 6, device="cuda", dtype=torch.int32)
        return lambda: K.check(L.aslp_splice_fwd(stream(), p(o), Ds * NO, p(x), Ds, Rs, Ds, p(off), NO)), (x, o, off)
    cs.append(Case("splice_fwd", "%dx%d->%d" % (Rs, Ds, Ds * NO), (1 + NO) * Ds * 4 * Rs, (1 + NO) * Ds * 4 * Rs, mk_sp))

    def mk_spb(i):
        e = torch.randn(Rs, Ds * NO, **f32); d = torch.empty(Rs, Ds, **f32)
        off = torch.arange(-5, 6, device="cuda", dtype=torch.int32)
        return lambda: K.check(L.aslp_splice_bwd(stream(), p(d), Ds, p(e), Ds * NO, Rs, Ds, p(off), NO)), (e, d, off)
    cs.append(Case("splice_bwd", "%dx%d<-%d" % (Rs, Ds, Ds * NO), (1 + NO) * Ds * 4 * Rs, (1 + NO) * Ds * 4 * Rs, mk_spb))

    # ---- FSMN memory block, cfg4: D = 512, 20/20 taps; T = 1000 (the cfg4 utterance) and T = 3000 (the component's limit)
    for T in (1000, 3000):
        Df, Pa, Fu = 512, 20, 20
        def mk_ff(i, T=T):
            x = torch.randn(T, Df, **f32); o = torch.empty_like(x); c = torch.randn(Pa + Fu + 1, Df, **f32) * 0.1
            return lambda: K.check(L.aslp_fsmn_fwd(stream(), p(o), Df, p(x), Df, T, Df, p(c), Df, Pa, Fu)), (x, o, c)
        cs.append(Case("fsmn_fwd", "T=%d D=%d 41 taps" % (T, Df), 2 * T * Df * 4 + 41 * Df * 4, 2 * T * Df * 4, mk_ff))

        def mk_fb(i, T=T):
            e = torch.randn(T, Df, **f32); d = torch.empty_like(e); c = torch.randn(Pa + Fu + 1, Df, **f32) * 0.1
            return lambda: K.check(L.aslp_fsmn_bwd(stream(), p(d), Df, p(e), Df, T, Df, p(c), Df, Pa, Fu)), (e, d, c)
        cs.append(Case("fsmn_bwd", "T=%d D=%d 41 taps" % (T, Df), 2 * T * Df * 4 + 41 * Df * 4, 2 * T * Df * 4, mk_fb))

        def mk_fg(i, T=T):
            x = torch.randn(T, Df, **f32); e = torch.randn(T, Df, **f32); g = torch.empty(Pa + Fu + 1, Df, **f32)
            return lambda: K.check(L.aslp_fsmn_coef_grad(stream(), p(g), Df, p(x), Df, p(e), Df, T, Df, Pa, Fu, 5.0)), (x, e, g)
        cs.append(Case("fsmn_coef_grad", "T=%d D=%d 41 taps" % (T, Df), 2 * T * Df * 4 + 41 * Df * 4, 2 * T * Df * 4, mk_fg))

    # ---- warp-ctc forward-backward, cfg3 utterances (T = 1000, K = 72, L = 100): 16 (one minibatch) and 2048 utterances
    for mb in (16, 2048):
        T, Kc, Lab = 1000, 72, 100
        def mk_ctc(i, mb=mb):
            rng = np.random.default_rng(i)
            acts = torch.randn(T, mb, Kc, **f32); grads = torch.zeros_like(acts)
            flat = np.ascontiguousarray(rng.integers(1, Kc, size=mb * Lab), np.int32)
            llen = np.full(mb, Lab, np.int32); ilen = np.full(mb, T, np.int32); costs = np.zeros(mb, np.float32)
            info = K.CtcComputeInfo(1, torch.cuda.current_stream().cuda_stream)
            size = ctypes.c_size_t(0)
            L.get_workspace_size(llen.ctypes.data, ilen.ctypes.data, Kc, mb, info, ctypes.addressof(size))
            ws = torch.empty(size.value + 256, dtype=torch.uint8, device="cuda")

            def fn():
                rc = L.compute_ctc_loss(p(acts), p(grads), flat.ctypes.data, llen.ctypes.data, ilen.ctypes.data, Kc, mb,
                                        costs.ctypes.data, p(ws), info)
                assert rc == 0
            return fn, (acts, grads, flat, llen, ilen, costs, ws)
        per_utt = 2 * T * Kc * 4 + 2 * T * (2 * Lab + 1) * 4
        cs.append(Case("ctc_fwd_bwd", "%d utts T=%d K=%d L=%d" % (mb, T, Kc, Lab), per_utt * mb, (2 * T * Kc * 4 + 3 * T * (2 * Lab + 1) * 4) * mb, mk_ctc))
    return cs


def run_case(c, reps_target_ms=60.0, once=False):
    nsets = 1 if once else int(min(64, max(2, np.ceil(2.2 * L2_BYTES / max(c.set_bytes, 1)))))
    fns, keep = [], []
    for i in range(nsets):
        fn, k = c.make(i)
        fns.append(fn); keep.append(k)
    if once:
        fns[0](); torch.cuda.synchronize()
        return {"kernel": c.name, "geometry": c.geom, "once": True}
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fns[0](); b.record(); torch.cuda.synchronize()
    est = max(a.elapsed_time(b), 1e-3)
    reps = int(min(2000, max(nsets, reps_target_ms / est)))
    reps = (reps + nsets - 1) // nsets * nsets
    best = None
    for _ in range(3):
        a.record()
        for r in range(reps):
            fns[r % nsets]()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        best = ms if best is None else min(best, ms)
    gbs = c.alg_bytes / best / 1e6
    return {"kernel": c.name, "geometry": c.geom, "us_per_launch": round(best * 1e3, 2), "algorithmic_MB": round(c.alg_bytes / 1e6, 3),
            "achieved_GBps": round(gbs, 1), "peak_GBps": PEAK, "frac": round(gbs / PEAK, 3), "buffer_sets": nsets, "reps": reps}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--once", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    PEAK = peak_gbs()
    torch.zeros(1, device="cuda")
    for c in cases():
        if args.once is not None and args.once not in c.name:
            continue
        if args.only is not None and args.only not in c.name:
            continue
        try:
            r = run_case(c, once=args.once is not None)
        except Exception as e:  # noqa: BLE001
            r = {"kernel": c.name, "geometry": c.geom, "error": str(e)[:300]}
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
