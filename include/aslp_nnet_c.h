/* aslp_nnet_c.h -- C handle API over the C++ host layer (libaslp_nnet.so): what a non-C++ caller (ctypes in the
 * parity tests and bench.py, or a trainer written in another language) binds.  Each function mirrors one method
 * of the reference's public C++ interface and cites it; errors are returned as non-zero status with the text in
 * aslp_nnet_last_error() (the C++ layer throws std::runtime_error like KALDI_ERR, src/base/kaldi-error.cc:146).
 * Host matrices are dense row-major fp32 (stride == cols); rows of recurrent nets are stream-interleaved (t*S+s).
 */
#ifndef ASLP_NNET_C_H_
#define ASLP_NNET_C_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* aslp_nnet_t;     /* kaldi::aslp_nnet::Nnet          (src/aslp-nnet/nnet-nnet.h:38-193)  */
typedef void* aslp_xent_t;     /* kaldi::aslp_nnet::Xent          (src/aslp-nnet/nnet-loss.h)          */
typedef void* aslp_warpctc_t;  /* kaldi::aslp_nnet::WarpCtc       (src/aslp-nnet/warp-ctc.h:27-129)    */
typedef void* aslp_worker_t;   /* kaldi::IWorker                  (src/aslp-parallel/itf.h:27-36)      */
typedef void* aslp_server_t;   /* kaldi::IServer                  (src/aslp-parallel/itf.h:38-43)      */

const char* aslp_nnet_last_error(void);
int aslp_nnet_select_device(int dev);                 /* CuDevice::Instantiate().SelectGpuId (aslp-nnet-train-*.cc) */
int aslp_nnet_srand(int seed);                        /* std::srand(seed) before Nnet::Init (aslp-nnet-init.cc:55) */
int aslp_nnet_set_gemm_precision(int precision);      /* ASLP_GEMM_3XTF32 (default) / ASLP_GEMM_TF32 / ASLP_GEMM_FP32 */
int aslp_nnet_device_sync(void);
unsigned long long aslp_nnet_launch_count(void);
/* training steps of aslp_train_step_xent that ran as a replayed recording (host/nnet-train-step.h) since the library was loaded */
long long aslp_nnet_step_replays(void);
/* CUDA events on the host layer's compute stream (the stream every kernel of this library is launched on) */
int aslp_nnet_event_record(int slot);                 /* slot in [0, 16) */
int aslp_nnet_event_elapsed_ms(int slot_a, int slot_b, float* ms);   /* synchronises on slot_b */
int aslp_nnet_pinned_alloc(void** host_ptr, size_t bytes);
int aslp_nnet_pinned_free(void* host_ptr);

int aslp_nnet_init(const char* proto_file, aslp_nnet_t* out);            /* Nnet::Init  (nnet-nnet.cc:561-603) */
int aslp_nnet_read(const char* model_file, aslp_nnet_t* out);            /* Nnet::Read  (nnet-nnet.cc:606-636) */
int aslp_nnet_write(aslp_nnet_t n, const char* file, int binary);        /* Nnet::Write (nnet-nnet.cc:638-653) */
int aslp_nnet_destroy(aslp_nnet_t n);
int aslp_nnet_input_dim(aslp_nnet_t n, int* dim);
int aslp_nnet_output_dim(aslp_nnet_t n, int* dim);
int aslp_nnet_num_components(aslp_nnet_t n, int* count);
int aslp_nnet_num_params(aslp_nnet_t n, int* count);
int aslp_nnet_info(aslp_nnet_t n, char* buf, size_t buf_bytes);          /* Nnet::Info */
int aslp_nnet_get_params(aslp_nnet_t n, float* host_out, int count);     /* Nnet::GetParams (nnet-nnet.cc:297-312) */
int aslp_nnet_set_train_options(aslp_nnet_t n, float learn_rate, float momentum, float l2_penalty, float l1_penalty);  /* SetTrainOptions */
int aslp_nnet_set_seq_lengths(aslp_nnet_t n, const int* lengths, int count);   /* Nnet::SetSeqLengths   (nnet-nnet.cc:492-523) */
int aslp_nnet_reset_streams(aslp_nnet_t n, const int* flags, int count);       /* Nnet::ResetLstmStreams (nnet-nnet.cc:467-490) */
int aslp_nnet_set_chunk_size(aslp_nnet_t n, int chunk_size);                   /* Nnet::SetChunkSize    (nnet-nnet.cc:525-532) */

/* Nnet::Propagate / Backpropagate / Feedforward with HOST buffers: the H2D copy of the input and the D2H copy of the
 * result (when the out pointer is non-NULL) are part of the call, as in the trainers (CuMatrix(feat) then CopyToMat). */
int aslp_nnet_propagate(aslp_nnet_t n, const float* host_in, int rows, int cols, float* host_out);
int aslp_nnet_feedforward(aslp_nnet_t n, const float* host_in, int rows, int cols, float* host_out);
int aslp_nnet_backpropagate(aslp_nnet_t n, const float* host_out_diff, int rows, int cols, float* host_in_diff);
/* per-component buffers of the last pass, for layer-by-layer parity checks */
int aslp_nnet_component_output(aslp_nnet_t n, int component, float* host_out, int rows, int cols);
int aslp_nnet_component_out_diff(aslp_nnet_t n, int component, float* host_out, int rows, int cols);

/* ---- losses ---- */
int aslp_xent_create(aslp_xent_t* out);
/* any frame-level objective behind the same handle: "xent" | "mse" | "multitask,<type>,<dim>,<weight>,..." (LossItf, nnet-loss.h:33-222;
 * the --objective-function values of the trainer mains) */
int aslp_loss_create(const char* objective, aslp_xent_t* out);
int aslp_xent_destroy(aslp_xent_t x);
int aslp_xent_report(aslp_xent_t x, char* buf, size_t buf_bytes, double stats5[5]);   /* Xent::Report; stats: avg-loss-numerator.. see .cc */
int aslp_warpctc_create(aslp_warpctc_t* out);
int aslp_warpctc_destroy(aslp_warpctc_t c);
int aslp_warpctc_report(aslp_warpctc_t c, char* buf, size_t buf_bytes);
int aslp_warpctc_rejected(aslp_warpctc_t c, int* n);   /* utterances rejected so far by the loss guard (StatAndAverageLossCheck, warp-ctc.cc:288-365) */

/* ---- one training minibatch, the loop body of the reference trainers ----
 * frame CE (aslp-nnet-train-frame.cc:110-124, -lstm-streams, -blstm-streams-lc):
 *   Propagate -> Xent::Eval(frame_mask, out, posterior) -> Backpropagate (which updates).
 * targets: one pdf-id per row; frame_mask may be NULL (all ones).  features_on_device != 0: `features` is a DEVICE
 * pointer with row stride cols rounded up to 4 floats (kernel-only timing); otherwise a host pointer. */
int aslp_train_step_xent(aslp_nnet_t n, aslp_xent_t x, const float* features, int features_on_device, int rows, int cols,
                         const int* targets, const float* frame_mask);
/* CTC (aslp-nnet-train-warp-ctc-streams.cc:175-198): SetSeqLengths -> learn_rate = norm_lr / valid frames (:177) when
 * norm_learn_rate > 0 -> Propagate -> WarpCtc::Eval -> ErrorRate (if with_error_rate) -> Backpropagate.
 * costs_out (host, nseq floats) receives -log p(z|x) per utterance. */
int aslp_train_step_ctc(aslp_nnet_t n, aslp_warpctc_t c, const float* features, int features_on_device, int rows, int cols,
                        const int* frame_num_utt, int nseq, const int* flat_labels, const int* label_lengths,
                        float norm_learn_rate, int with_error_rate, float* costs_out);
/* Eesen-style CTC (kaldi::aslp_nnet::Ctc, src/aslp-nnet/ctc-loss.{h,cc}); the loop body of aslp-nnet-train-ctc-streams.cc:160-205:
 * SetSeqLengths, learn-rate normalisation, Propagate, Ctc::EvalParallel, ErrorRateMSeq (optional), Backpropagate.
 * obj_out[nseq] receives -log p(z|x) per sequence. */
typedef void* aslp_eesenctc_t;
int aslp_eesenctc_create(aslp_eesenctc_t* out);
int aslp_eesenctc_destroy(aslp_eesenctc_t c);
int aslp_eesenctc_report(aslp_eesenctc_t c, char* buf, size_t buf_bytes);
int aslp_train_step_ctc_eesen(aslp_nnet_t n, aslp_eesenctc_t c, const float* features, int features_on_device, int rows, int cols,
                              const int* frame_num_utt, int nseq, const int* flat_labels, const int* label_lengths,
                              float norm_learn_rate, int with_error_rate, float* obj_out);
/* pinned host staging + device upload helpers for the bench (aslp_malloc_host / CuMatrix with padded stride) */
int aslp_nnet_upload(const float* host, int rows, int cols, float** device_out, int* stride_out);
int aslp_nnet_free_device(float* device_ptr);

/* ---- aslp-parallel workers (src/aslp-parallel/itf.h:27-36, bsp-worker.cc, bmuf-worker.cc, sod-worker.cc) ----
 * type: "bsp" | "bmuf" | "sod".  nccl_id: the 128-byte id from aslp_comm_unique_id() of rank 0, shipped by the launcher. */
int aslp_worker_create(const char* type, const char nccl_id[128], int nranks, int rank, float bmuf_momentum, float bmuf_learn_rate,
                       const char* sod_solver, aslp_worker_t* out);
int aslp_worker_init_param(aslp_worker_t w, aslp_nnet_t n);                       /* IWorker::InitParam(GetGpuParams) */
int aslp_worker_synchronize(aslp_worker_t w, int num_frames, int* keep_going);    /* IWorker::Synchronize */
/* The exchange pipelined by layer (not in the reference, same arithmetic per tensor): register the tensors with
 * aslp_worker_init_param_by_component on EVERY rank (synchronize then exchanges component by component, top layer first); a
 * trainer that knows a synchronisation is due after the coming minibatch calls begin_synchronize(frames) before the minibatch
 * and end_synchronize after it instead of synchronize(frames): each component's tensors are exchanged on the worker's own stream as
 * soon as its Update is enqueued, while the layers below still back-propagate.  bsp, bmuf and sod (can_overlap). */
int aslp_worker_init_param_by_component(aslp_worker_t w, aslp_nnet_t n);
int aslp_worker_can_overlap(aslp_worker_t w, int* yes);
int aslp_worker_begin_synchronize(aslp_worker_t w, int num_frames);
int aslp_worker_end_synchronize(aslp_worker_t w, int* keep_going);
int aslp_worker_stop(aslp_worker_t w);                                            /* IWorker::Stop */
int aslp_worker_reduce_acc_stat(aslp_worker_t w, aslp_nnet_t n);                  /* MpiNode::ReduceAccStat (mpi-node.h:76-91) */
int aslp_worker_destroy(aslp_worker_t w);
/* worker types also: "easgd" (bmuf_learn_rate carries alpha, easgd-worker.h:20), "asgd" / "masgd" (asgd-worker.cc); rank 0 is
 * the server of these modes and creates an aslp_server_t instead of a worker.
 * ---- async parameter servers (easgd-server.cc, asgd-server.cc, masgd-server.cc): rank 0, Run() returns when every worker
 * has sent kMsgFinished.  The any-source message channel is loopback TCP on ASLP_CTRL_PORT (default MASTER_PORT + 1). */
int aslp_server_create(const char* type, const char nccl_id[128], int nranks, float alpha, int sync_period, float momentum, aslp_server_t* out);
int aslp_server_init_param(aslp_server_t s, aslp_nnet_t n);                       /* IServer::InitParam(GetGpuParams) */
int aslp_server_run(aslp_server_t s);                                             /* IServer::Run */
int aslp_server_destroy(aslp_server_t s);
/* the control channel alone (host only; what replaces MPI_Recv(MPI_ANY_SOURCE, kTagMsg), itf.h:13-22) */
int aslp_ctrl_server_create(int port, int nworkers, void** out);
int aslp_ctrl_server_recv_any(void* server, int* worker_rank, int* msg_type);
int aslp_ctrl_server_destroy(void* server);
int aslp_ctrl_client_create(int port, int rank, void** out);
int aslp_ctrl_client_send(void* client, int msg_type);
int aslp_ctrl_client_destroy(void* client);

#ifdef __cplusplus
}
#endif
#endif
