"""ctypes mirror of the host C API (include/aslp_nnet_c.h -> libaslp_nnet.so).  Thin: every method is one C call into
the C++ Nnet / Xent / WarpCtc / IWorker mirror; numpy arrays are passed as host buffers.  No computation here."""
import ctypes

import numpy as np

from . import cuda_lib, host_lib

P = ctypes.c_void_p


def _ck(rc):
    if rc != 0:
        raise RuntimeError(host_lib().aslp_nnet_last_error().decode(errors="replace"))


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def select_device(dev):
    _ck(host_lib().aslp_nnet_select_device(int(dev)))


def srand(seed):
    host_lib().aslp_nnet_srand(int(seed))


def set_gemm_precision(p):
    host_lib().aslp_nnet_set_gemm_precision(int(p))


def device_sync():
    _ck(host_lib().aslp_nnet_device_sync())


def launch_count():
    return int(host_lib().aslp_nnet_launch_count())


def step_replays():
    """train_step_xent calls that ran as a replayed recording (CuStepGraph) so far"""
    return int(host_lib().aslp_nnet_step_replays())


class Nnet:
    """kaldi::aslp_nnet::Nnet (src/aslp-nnet/nnet-nnet.h:38-193)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def init(cls, proto_file):
        h = P()
        _ck(host_lib().aslp_nnet_init(str(proto_file).encode(), ctypes.byref(h)))
        return cls(h)

    @classmethod
    def read(cls, model_file):
        h = P()
        _ck(host_lib().aslp_nnet_read(str(model_file).encode(), ctypes.byref(h)))
        return cls(h)

    def write(self, path, binary=True):
        _ck(host_lib().aslp_nnet_write(self.h, str(path).encode(), int(binary)))

    def close(self):
        if self.h:
            host_lib().aslp_nnet_destroy(self.h)
            self.h = None

    def _int(self, fn):
        v = ctypes.c_int(0)
        _ck(fn(self.h, ctypes.byref(v)))
        return v.value

    @property
    def input_dim(self):
        return self._int(host_lib().aslp_nnet_input_dim)

    @property
    def output_dim(self):
        return self._int(host_lib().aslp_nnet_output_dim)

    @property
    def num_components(self):
        return self._int(host_lib().aslp_nnet_num_components)

    @property
    def num_params(self):
        return self._int(host_lib().aslp_nnet_num_params)

    def info(self):
        buf = ctypes.create_string_buffer(1 << 16)
        _ck(host_lib().aslp_nnet_info(self.h, buf, len(buf)))
        return buf.value.decode()

    def get_params(self):
        out = np.zeros(self.num_params, np.float32)
        _ck(host_lib().aslp_nnet_get_params(self.h, out.ctypes.data, out.size))
        return out

    def set_train_options(self, learn_rate=0.008, momentum=0.0, l2_penalty=0.0, l1_penalty=0.0):
        _ck(host_lib().aslp_nnet_set_train_options(self.h, learn_rate, momentum, l2_penalty, l1_penalty))

    def set_seq_lengths(self, lengths):
        a = _i32(lengths)
        _ck(host_lib().aslp_nnet_set_seq_lengths(self.h, a.ctypes.data, a.size))

    def reset_streams(self, flags):
        a = _i32(flags)
        _ck(host_lib().aslp_nnet_reset_streams(self.h, a.ctypes.data, a.size))

    def set_chunk_size(self, n):
        _ck(host_lib().aslp_nnet_set_chunk_size(self.h, int(n)))

    def propagate(self, x):
        x = _f32(x)
        out = np.zeros((x.shape[0], self.output_dim), np.float32)
        _ck(host_lib().aslp_nnet_propagate(self.h, x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data))
        return out

    def feedforward(self, x):
        x = _f32(x)
        out = np.zeros((x.shape[0], self.output_dim), np.float32)
        _ck(host_lib().aslp_nnet_feedforward(self.h, x.ctypes.data, x.shape[0], x.shape[1], out.ctypes.data))
        return out

    def backpropagate(self, out_diff):
        d = _f32(out_diff)
        ind = np.zeros((d.shape[0], self.input_dim), np.float32)
        _ck(host_lib().aslp_nnet_backpropagate(self.h, d.ctypes.data, d.shape[0], d.shape[1], ind.ctypes.data))
        return ind

    def component_output(self, c, rows, cols):
        out = np.zeros((rows, cols), np.float32)
        _ck(host_lib().aslp_nnet_component_output(self.h, c, out.ctypes.data, rows, cols))
        return out

    def component_out_diff(self, c, rows, cols):
        out = np.zeros((rows, cols), np.float32)
        _ck(host_lib().aslp_nnet_component_out_diff(self.h, c, out.ctypes.data, rows, cols))
        return out


class Xent:
    """a frame-level objective (LossItf): Xent by default, Loss("mse") / Loss("multitask,...") for the others"""
    def __init__(self, objective="xent"):
        self.h = P()
        _ck(host_lib().aslp_loss_create(objective.encode(), ctypes.byref(self.h)))

    def report(self):
        buf = ctypes.create_string_buffer(4096)
        st = (ctypes.c_double * 5)()
        _ck(host_lib().aslp_xent_report(self.h, buf, len(buf), st))
        return buf.value.decode(), list(st)


Loss = Xent


class WarpCtc:
    def __init__(self):
        self.h = P()
        _ck(host_lib().aslp_warpctc_create(ctypes.byref(self.h)))

    def report(self):
        buf = ctypes.create_string_buffer(4096)
        _ck(host_lib().aslp_warpctc_report(self.h, buf, len(buf)))
        return buf.value.decode()


def warpctc_rejected(ctc):
    """utterances whose derivative the 6-sigma / (0, 3000) loss guard has zeroed so far"""
    n = ctypes.c_int(0)
    _ck(host_lib().aslp_warpctc_rejected(ctc.h, ctypes.byref(n)))
    return n.value


class EesenCtc:
    """kaldi::aslp_nnet::Ctc (src/aslp-nnet/ctc-loss.h): CTC on the softmax outputs, error back-propagated through the softmax."""

    def __init__(self):
        self.h = P()
        _ck(host_lib().aslp_eesenctc_create(ctypes.byref(self.h)))

    def report(self):
        buf = ctypes.create_string_buffer(4096)
        _ck(host_lib().aslp_eesenctc_report(self.h, buf, len(buf)))
        return buf.value.decode()


def train_step_ctc_eesen(nnet, ctc, feats, frame_num_utt, labels, norm_learn_rate=0.0, with_error_rate=False):
    """Loop body of aslp-nnet-train-ctc-streams.cc.  Returns -log p(z|x) per sequence."""
    lens = _i32(frame_num_utt)
    fl = _i32(np.concatenate([np.asarray(l, np.int32) for l in labels]))
    ll = _i32([len(l) for l in labels])
    obj = np.zeros(lens.size, np.float32)
    feats = _f32(feats)
    rows, cols = feats.shape
    _ck(host_lib().aslp_train_step_ctc_eesen(nnet.h, ctc.h, P(feats.ctypes.data), 0, rows, cols, lens.ctypes.data, lens.size, fl.ctypes.data,
                                             ll.ctypes.data, float(norm_learn_rate), int(with_error_rate), obj.ctypes.data))
    return obj


def train_step_xent(nnet, xent, feats, targets, frame_mask=None, on_device=False, rows=None, cols=None):
    """Loop body of aslp-nnet-train-frame / -lstm-streams / -blstm-streams-lc (Propagate, Xent::Eval, Backpropagate)."""
    t = _i32(targets)
    m = _f32(frame_mask) if frame_mask is not None else None
    if on_device:
        ptr = P(feats)
    else:
        feats = _f32(feats)
        rows, cols = feats.shape
        ptr = P(feats.ctypes.data)
    _ck(host_lib().aslp_train_step_xent(nnet.h, xent.h, ptr, int(on_device), rows, cols, t.ctypes.data,
                                        m.ctypes.data if m is not None else None))


def train_step_ctc(nnet, ctc, feats, frame_num_utt, labels, norm_learn_rate=0.0, with_error_rate=False, on_device=False,
                   rows=None, cols=None, flat=None):
    """Loop body of aslp-nnet-train-warp-ctc-streams.cc:175-198.  Returns the per-utterance costs."""
    lens = _i32(frame_num_utt)
    if flat is None:
        flat = (_i32(np.concatenate([np.asarray(l, np.int32) for l in labels])), _i32([len(l) for l in labels]))
    fl, ll = flat
    costs = np.zeros(lens.size, np.float32)
    if on_device:
        ptr = P(feats)
    else:
        feats = _f32(feats)
        rows, cols = feats.shape
        ptr = P(feats.ctypes.data)
    _ck(host_lib().aslp_train_step_ctc(nnet.h, ctc.h, ptr, int(on_device), rows, cols, lens.ctypes.data, lens.size, fl.ctypes.data,
                                       ll.ctypes.data, norm_learn_rate, int(with_error_rate), costs.ctypes.data))
    return costs


def upload(arr):
    """Copy a host matrix into a device buffer with the padded row stride the C-ABI expects; returns (ptr, stride)."""
    a = _f32(arr)
    ptr = P()
    stride = ctypes.c_int(0)
    _ck(host_lib().aslp_nnet_upload(a.ctypes.data, a.shape[0], a.shape[1], ctypes.byref(ptr), ctypes.byref(stride)))
    return ptr.value, stride.value


class Worker:
    """kaldi::IWorker (src/aslp-parallel/itf.h:27-36): bsp | bmuf | sod over NCCL."""

    def __init__(self, kind, nccl_id, nranks, rank, bmuf_momentum=0.9, bmuf_learn_rate=1.0, sod_solver="momentum"):
        self.h = P()
        idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128)
        _ck(host_lib().aslp_worker_create(kind.encode(), idbuf, nranks, rank, bmuf_momentum, bmuf_learn_rate, sod_solver.encode(),
                                          ctypes.byref(self.h)))

    def init_param(self, nnet):
        _ck(host_lib().aslp_worker_init_param(self.h, nnet.h))

    def init_param_by_component(self, nnet):
        """As init_param, remembering which tensors belong to which component: every synchronisation then goes component by
        component (all ranks must register the same way), and begin_synchronize / end_synchronize can pipeline it by layer."""
        _ck(host_lib().aslp_worker_init_param_by_component(self.h, nnet.h))

    def can_overlap(self):
        y = ctypes.c_int(0)
        _ck(host_lib().aslp_worker_can_overlap(self.h, ctypes.byref(y)))
        return bool(y.value)

    def synchronize(self, num_frames):
        k = ctypes.c_int(0)
        _ck(host_lib().aslp_worker_synchronize(self.h, int(num_frames), ctypes.byref(k)))
        return bool(k.value)

    def begin_synchronize(self, num_frames):
        """Before the minibatch after which synchronize(num_frames) would have been called (IWorker::BeginSynchronize)."""
        _ck(host_lib().aslp_worker_begin_synchronize(self.h, int(num_frames)))

    def end_synchronize(self):
        k = ctypes.c_int(0)
        _ck(host_lib().aslp_worker_end_synchronize(self.h, ctypes.byref(k)))
        return bool(k.value)

    def stop(self):
        _ck(host_lib().aslp_worker_stop(self.h))

    def close(self):
        if self.h:
            host_lib().aslp_worker_destroy(self.h)
            self.h = None


class Server:
    """kaldi::IServer (src/aslp-parallel/itf.h:38-43): easgd | asgd | masgd parameter server on rank 0."""

    def __init__(self, kind, nccl_id, nranks, alpha=0.5, sync_period=1000, momentum=0.9):
        self.h = P()
        idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128)
        _ck(host_lib().aslp_server_create(kind.encode(), idbuf, nranks, alpha, sync_period, momentum, ctypes.byref(self.h)))

    def init_param(self, nnet):
        _ck(host_lib().aslp_server_init_param(self.h, nnet.h))

    def run(self):
        _ck(host_lib().aslp_server_run(self.h))

    def close(self):
        if self.h:
            host_lib().aslp_server_destroy(self.h)
            self.h = None


def nccl_unique_id():
    buf = ctypes.create_string_buffer(128)
    rc = cuda_lib().aslp_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError(cuda_lib().aslp_last_error().decode())
    return buf.raw
