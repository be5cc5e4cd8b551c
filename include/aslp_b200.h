/* aslp_b200.h -- C-ABI of libaslp_b200.so: the sm_100a backing for the aslp-nnet
 * Component Propagate / Backpropagate / Update path and the aslp-parallel averaging.
 *
 * It replaces, for this path only, what src/aslp-cudamatrix offers the reference:
 * the 257 `extern "C" cudaF_*` launchers (src/aslp-cudamatrix/cu-kernels-ansi.h:32-401),
 * cuBLAS behind CuMatrixBase::AddMatMat (src/aslp-cudamatrix/cu-matrix.cc:1027-1062,
 * cublas-wrappers.h:28-38) and CuDevice::Malloc/Free (cu-device.h:54-59) -- re-cut
 * coarser: ONE symbol per fused operation.
 *
 * Conventions (all functions):
 *   - plain C symbols, raw DEVICE pointers + {rows, cols, stride} in ELEMENTS (like
 *     MatrixDim, src/aslp-cudamatrix/cu-matrixdim.h:51-55), scalars by value;
 *   - row-major fp32 matrices; row stride (ld*) must be a multiple of 4 floats and base
 *     pointers 16-byte aligned (what CuMatrix's pitched allocation gives the reference);
 *   - every call takes an explicit stream (a cudaStream_t passed as void*), is
 *     asynchronous, never allocates behind the caller's back (workspaces are explicit),
 *     returns 0 on success like ctcStatus_t (src/warp-ctc/include/ctc.h:16-22) and
 *     never throws; aslp_last_error() gives the text for a non-zero status;
 *   - there is NO CPU path: without a CUDA device every compute call fails.
 */
#ifndef ASLP_B200_H_
#define ASLP_B200_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* aslp_stream_t; /* cudaStream_t */

enum {
  ASLP_STATUS_SUCCESS = 0,
  ASLP_STATUS_MEMOPS_FAILED = 1,
  ASLP_STATUS_INVALID_VALUE = 2,
  ASLP_STATUS_EXECUTION_FAILED = 3,
  ASLP_STATUS_UNKNOWN_ERROR = 4
};

/* ---- runtime: device, memory, streams (replaces CuDevice, cu-device.h:40-170) ---- */
const char* aslp_last_error(void);
unsigned long long aslp_launch_count(void);       /* kernels launched by this library so far */
int aslp_device_count(int* n);
int aslp_set_device(int dev);                     /* CuDevice::SelectGpuId */
int aslp_get_device(int* dev);                    /* the calling thread's current device (helper threads inherit it explicitly) */
int aslp_malloc(void** dptr, size_t bytes);       /* CuDevice::Malloc (cu-device.h:54) */
int aslp_free(void* dptr);                        /* CuDevice::Free   (cu-device.h:59) */
int aslp_malloc_host(void** hptr, size_t bytes);  /* pinned staging for CopyFromMat/CopyToMat */
int aslp_free_host(void* hptr);
int aslp_memset(aslp_stream_t s, void* dptr, int value, size_t bytes);
int aslp_memset2d(aslp_stream_t s, void* dptr, size_t pitch, int value, size_t width, size_t height);   /* strided views, bytes */
int aslp_memcpy_h2d(aslp_stream_t s, void* dst, const void* src, size_t bytes);
int aslp_memcpy_d2h(aslp_stream_t s, void* dst, const void* src, size_t bytes);
int aslp_memcpy_d2d(aslp_stream_t s, void* dst, const void* src, size_t bytes);
/* strided 2-D copies, widths/pitches in BYTES (CuMatrix::CopyFromMat, cu-matrix.cc:236-330) */
int aslp_memcpy2d_h2d(aslp_stream_t s, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height);
int aslp_memcpy2d_d2h(aslp_stream_t s, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height);
int aslp_memcpy2d_d2d(aslp_stream_t s, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height);
int aslp_stream_create(aslp_stream_t* s);
int aslp_stream_destroy(aslp_stream_t s);
int aslp_stream_sync(aslp_stream_t s);
int aslp_device_sync(void);
/* CUDA events: *event is created on first use; elapsed synchronises on `b` */
int aslp_event_record(aslp_stream_t s, void** event);
int aslp_event_elapsed_ms(void* a, void* b, float* ms);
int aslp_stream_wait_event(aslp_stream_t s, void* event);   /* later work on s starts after `event` (recorded on another stream) */
int aslp_event_destroy(void* event);                        /* NULL is fine */
int aslp_scratch_release(aslp_stream_t s);                  /* frees the library's per-stream scratch arena before a stream is destroyed */
int aslp_event_sync(void* event);                           /* the host waits for `event` (a pinned staging slot is free again) */
/* Static-shape step replay.  The reference enqueues every kernel of every minibatch from the host and device-syncs after
 * most of them (CU_SAFE_CALL, src/aslp-cudamatrix/cu-common.h:38-45); at a 256-frame DNN minibatch (BASELINE cfg1) the
 * step is then bound by host enqueue time, not by the device.  A step whose shapes, pointers and scalars repeat is
 * captured once from the stream it is enqueued on and replayed as one executable graph.
 *   aslp_graph_begin(s)                   later work enqueued on s (and on streams that join it through events) is recorded, not run
 *   aslp_graph_end(s, &exec, &kernels)    ends the recording and instantiates it; non-zero (exec = NULL) when the recording
 *                                         was invalidated -- nothing ran, the caller enqueues the step again without recording
 *   aslp_graph_launch(exec, s), aslp_graph_destroy(exec)
 *   aslp_alloc_epoch()                    changes whenever device memory has been released (aslp_free, scratch regrowth): a
 *                                         recording is only valid while the epoch it was made in lasts
 *   aslp_count_launches(n)                adds the kernels of a replay to the launch counter (aslp_launch_count) */
int aslp_graph_begin(aslp_stream_t s);
int aslp_graph_end(aslp_stream_t s, void** exec, int* kernel_nodes);
int aslp_graph_launch(void* exec, aslp_stream_t s);
int aslp_graph_destroy(void* exec);
unsigned long long aslp_alloc_epoch(void);
int aslp_count_launches(unsigned long long n);

/* ---- dense contraction: CuMatrixBase::AddMatMat (cu-matrix.cc:1027-1062) ----
 * C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] + beta * C  (+ bias[n] broadcast over rows)
 * then, if clip > 0, C = min(max(C, -clip), clip)  (ApplyFloor/ApplyCeiling on *_corr_,
 * src/aslp-nnet/nnet-blstm-projected-streams-lc.h:1002-1017 fused into the epilogue).
 * trans_a: A is stored [K,M]; trans_b: B is stored [N,K] (Kaldi kTrans).
 * precision: ASLP_GEMM_3XTF32 = fp32-grade (default): chunk-sized products run as ASLP_GEMM_F16X3, smaller ones split tf32 hi / lo in the main loop
 *                               (3 tcgen05 MMAs per k-step either way),
 *            ASLP_GEMM_F16X3  = fp32-grade: operands split once into fp16 hi / lo planes (row-wise power-of-two scaling), 3 kind::f16 passes,
 *            ASLP_GEMM_TF32   = single-pass TF32 (looser bound),
 *            ASLP_GEMM_FP32   = CUDA-core fp32 FMA (exact-order-free fp32; small/odd shapes).
 * workspace: used only for split-K; aslp_gemm_workspace_bytes() gives the size (may be 0). */
enum { ASLP_GEMM_3XTF32 = 0, ASLP_GEMM_TF32 = 1, ASLP_GEMM_FP32 = 2, ASLP_GEMM_F16X3 = 3 };
/* caps the persistent grid of the tensor-core GEMMs launched by the CALLING THREAD (0 = one CTA per SM): weight-gradient products
 * issued on a side stream leave SMs free for the persistent recurrence kernel they are meant to overlap with */
int aslp_gemm_set_cta_limit(int max_ctas);
size_t aslp_gemm_workspace_bytes(int M, int N, int K);
int aslp_gemm(aslp_stream_t s, int trans_a, int trans_b, int M, int N, int K,
              float alpha, const float* A, int lda, const float* B, int ldb,
              float beta, float* C, int ldc, const float* bias, float clip,
              int precision, void* workspace, size_t workspace_bytes);

/* aslp_gemm with what the components do to the product right afterwards folded in (SURVEY 2.4's fusion targets):
 *   act         0 none, 1 + ASLP_ACT_* : C = f(C)            Affine followed by Sigmoid / Tanh / ReLU (nnet-activation.h:153-199, 275-303)
 *   dact_y      C = f'(y) * C with y the activation's OUTPUT  the DiffSigmoid / DiffTanh / ReLU mask the activation's Backpropagate
 *                                                             applies to what the next layer's backward product delivers
 *   update_w    W -= update_lr * C after C = beta*C + alpha*AB the SGD apply behind the weight-gradient product, C = *_corr_
 *                                                             (nnet-affine-transform.h:210,237)
 * order: product, beta, bias, clip, act, dact, store, update.  Folded into the split-K reduction when the product takes it (the
 * few-tile shapes of 256- to 1000-frame minibatches, where every saved launch counts); otherwise the same steps run as
 * launches of their own behind the product, so the result does not depend on the path. */
typedef struct {
  int act;
  const float* dact_y; int dact_ldy; int dact_kind;      /* ASLP_ACT_* of the activation whose derivative is applied; y may not alias C */
  float* update_w; int update_ldw; float update_lr;
  int reduce_in_launch;   /* non-zero: a split-K product may reduce its partial tiles inside the launch -- the splits of a tile run as
                           * a thread-block cluster and add their tiles through distributed shared memory -- instead of in a
                           * second pass; the library falls back to the second pass when the clusters of the product do not all
                           * fit the device.  Same summation order, bit-identical results; at par with the second pass on B200
                           * (profiles/r02_gemm_in_launch_reduce.txt), the host layer leaves it 0. */
} aslp_gemm_epilogue_t;
/* diagnostic: clusters of `cluster_size` (2..8) CTAs of the split-K kernel the device runs at once; aslp_gemm_ex only reduces inside
 * the launch when all the clusters of a product fit */
int aslp_gemm_cluster_fit(int cluster_size);
int aslp_gemm_ex(aslp_stream_t s, int trans_a, int trans_b, int M, int N, int K,
                 float alpha, const float* A, int lda, const float* B, int ldb,
                 float beta, float* C, int ldc, const float* bias, float clip,
                 int precision, void* workspace, size_t workspace_bytes, const aslp_gemm_epilogue_t* epi);

/* ---- pointwise / reductions (cu-kernels.cu:1802-1858, 700-760, 416-437, 1346-1505) ---- */
enum { ASLP_ACT_SIGMOID = 0, ASLP_ACT_TANH = 1, ASLP_ACT_RELU = 2 };
/* Sigmoid/Tanh/ReLU::PropagateFnc (src/aslp-nnet/nnet-activation.h:153-199,276-298) */
int aslp_act_fwd(aslp_stream_t s, int kind, float* out, int ldo, const float* in, int ldi, int rows, int cols);
/* BackpropagateFnc: sigmoid y(1-y)e, tanh (1-y^2)e use the OUTPUT y; relu uses the INPUT x (x>0 ? e : 0) */
int aslp_act_bwd(aslp_stream_t s, int kind, float* in_diff, int ldd, const float* y_or_x, int ldy,
                 const float* out_diff, int lde, int rows, int cols);
/* CuMatrixBase::CopyRows (cu-matrix.cc:2055-2080; kernel cu-kernels.cu:1090-1110): dst[r,:] = src[idx[r],:], idx < 0 -> zero row.
 * The frame shuffle of MatrixRandomizer::Randomize (src/aslp-nnet/nnet-randomizer.cc:75-90). */
int aslp_copy_rows(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, const int* idx_dev, int rows, int cols);
/* Softmax::PropagateFnc, ApplySoftMaxPerRow (nnet-activation.h:49-52; kaldi-vector.cc:852-859) */
int aslp_softmax_rows(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int cols);
/* dst = alpha*src + beta*dst elementwise (AddMat / CopyFromMat / Scale; cu-kernels.cu:584) */
int aslp_axpby(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int cols, float alpha, float beta);
/* dst[r,c] = alpha*vec[c] + beta*dst[r,c]  (AddVecToRows, cu-kernels.cu:700) */
int aslp_add_vec_to_rows(aslp_stream_t s, float* dst, int ldd, int rows, int cols, const float* vec, float alpha, float beta);
/* vec[c] = alpha * sum_r mat[r,c] + beta*vec[c], then optional clip (CuVector::AddRowSumMat, cu-vector.cc:1145-1166) */
int aslp_col_sum(aslp_stream_t s, float* vec, const float* mat, int ldm, int rows, int cols, float alpha, float beta, float clip);
/* bias gradient and bias update of an affine layer in one launch (nnet-affine-transform.h:211,237-238):
 * corr[c] = momentum * corr[c] + sum_r diff[r,c];  bias[c] -= lr * corr[c] */
int aslp_bias_grad_update(aslp_stream_t s, float* bias, float* corr, const float* diff, int ldd, int rows, int cols, float momentum, float lr);
/* vec[c] = alpha * sum_r a[r,c]*b[r,c] + beta*vec[c], optional clip (AddDiagMatMat(kTrans,kNoTrans), cu-kernels.cu:992; peephole grads) */
int aslp_col_dot(aslp_stream_t s, float* vec, const float* a, int lda, const float* b, int ldb, int rows, int cols, float alpha, float beta, float clip);
/* clamp to [lo, hi] (ApplyFloor + ApplyCeiling) */
int aslp_clamp(aslp_stream_t s, float* dst, int ldd, int rows, int cols, float lo, float hi);
/* cu::RegularizeL1 (src/aslp-cudamatrix/cu-math.cc:37-77) */
int aslp_regularize_l1(aslp_stream_t s, float* w, int ldw, float* grad, int ldg, int rows, int cols, float l1, float lr);
/* AffineTransform max-norm renormalisation of rows (nnet-affine-transform.h:232-243) */
int aslp_max_norm_rows(aslp_stream_t s, float* w, int ldw, int rows, int cols, float max_norm);
/* transpose: dst[c,r] = src[r,c] */
int aslp_transpose(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int cols);
/* sum of all elements / finite check (Nnet::Check, Xent asserts): out[0]=sum (double), out[1]=#non-finite */
int aslp_sum_check(aslp_stream_t s, const float* m, int ldm, int rows, int cols, double* out2_dev);

/* ---- Xent::Eval (src/aslp-nnet/nnet-loss.cc:63-156), one pass ----
 * sparse one-target-per-frame form: tgt_idx[r] in [0,cols), tgt_w[r] = posterior weight.
 * frame_w[r] = frame mask.  diff = (y - t) * (frame_w * sum_k t).
 * stats_dev[5] (double, ACCUMULATED into): cross-entropy, entropy, likelihood, correct, frames. */
int aslp_xent_sparse(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, int rows, int cols,
                     const int* tgt_idx, const float* tgt_w, const float* frame_w, double* stats_dev);
/* dense-target form (Posterior with several pdfs per frame, PosteriorToMatrix) */
int aslp_xent_dense(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt,
                    int rows, int cols, const float* frame_w, double* stats_dev);
/* Mse::Eval (src/aslp-nnet/nnet-loss.cc:205-258): diff = w (y - t), stats_dev[0] += 0.5 * sum w * diff^2 (the reference weights the
 * squared, already weighted difference once more; kept).  stats_dev: one device double, never reset by the call. */
int aslp_mse(aslp_stream_t s, float* diff, int ldd, const float* y, int ldy, const float* tgt, int ldt, int rows, int cols,
             const float* frame_w, double* stats_dev);
/* row arg-max (FindRowMaxId, cu-kernels.cu:2141): first maximal index */
int aslp_row_argmax(aslp_stream_t s, int* idx, const float* m, int ldm, int rows, int cols);

/* ---- forwarder post-processing (src/aslp-nnetbin/aslp-nnet-forward.cc:184-207; nnet-pdf-prior.cc:73-86), one pass ----
 * m = [log(m + log_add)] ; m[:,0] -= blank_shift (if > 0) ; m -= prior_scale * log_priors (if given).
 * stats5_dev (float[5], overwritten): min / max of the input, min / max before the prior stage, #non-finite outputs. */
int aslp_posterior_finalize(aslp_stream_t s, float* m, int ldm, int rows, int cols, int apply_log, float log_add, float blank_shift,
                            const float* log_priors_dev, float prior_scale, float* stats5_dev);

/* ---- Splice (src/aslp-nnet/nnet-various.h:139-175; cu-math.cc:153-166) ---- */
int aslp_splice_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int dim,
                    const int* offsets_dev, int n_offsets);
/* reference's backward: in_diff[t] = sum_c out_diff[clamp(t+off[c]), c-th block] (quirk kept) */
int aslp_splice_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* out_diff, int ldo, int rows, int dim,
                    const int* offsets_dev, int n_offsets);

/* ---- RowConvolution (src/aslp-nnet/nnet-row-convolution.cc:90-169) ----
 * stream-interleaved rows t*S+s; w is [dim, future+1]; seq_len_dev int[S]. */
int aslp_rowconv_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int T, int S, int dim,
                     const float* w, int ldw, int future, const int* seq_len_dev);
int aslp_rowconv_bwd(aslp_stream_t s, float* in_diff, int ldd, float* w_diff, int ldwd,
                     const float* in, int ldi, const float* out_diff, int ldo, int T, int S, int dim,
                     const float* w, int ldw, int future, const int* seq_len_dev);

/* ---- BatchNormalization (src/aslp-nnet/nnet-batch-normalization.h:139-284) ----
 * train fwd: batch mean / inv-std (var_floor), out = xhat*scale+shift, fp64 running sums acc_mean += sum x,
 * acc_var += sum x^2 (:217-219).  xhat may be NULL (the backward recomputes it from in, mean, inv_std and ignores its
 * xhat argument); when given it is filled for callers that want it. */
int aslp_bn_fwd_train(aslp_stream_t s, float* out, int ldo, float* xhat, int ldx, const float* in, int ldi,
                      int rows, int cols, const float* scale, const float* shift, float var_floor,
                      float* mean, float* inv_std, double* acc_mean, double* acc_var);
/* eval fwd with given mean / inv_std (FeedforwardFnc global-stats branch :167-174) */
int aslp_bn_fwd_eval(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int cols,
                     const float* scale, const float* shift, const float* mean, const float* inv_std);
/* bwd: dscale = mmt*dscale + sum xhat*dy; dshift = mmt*dshift + sum dy; in_diff per :236-277 */
int aslp_bn_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* xhat, int ldx,
                const float* out_diff, int ldo, int rows, int cols, const float* scale,
                const float* mean, const float* inv_std, float momentum, float* dscale, float* dshift);

/* ---- ConvolutionalComponent / MaxPoolingComponent (src/aslp-nnet/nnet-convolutional-component.h:263-421,
 * nnet-max-pooling-component.h:100-156) ----
 * gather : patches[(b*num_patches + p)*ldp + s*patch_dim + d] = in[b*ldi + p*patch_step + s*patch_stride + d]
 *          (replaces column_map + CopyCols; with this layout the per-patch AddMatMat loop is ONE aslp_gemm:
 *          [rows*num_patches, filter_dim] x filters^T -> out viewed as [rows*num_patches, num_filters]);
 *          when ldp is filter_dim rounded up to a multiple of 4 the rounding columns are written too (zeros), otherwise
 *          columns beyond filter_dim are left alone
 * scatter: in_diff[b, c] = sum over the patch positions that read column c, in ascending p (replaces ReverseIndexes /
 *          RearrangeIndexes / AddCols); in_diff is overwritten
 * maxpool fwd: out[b, q*S + j] = max(-1e20, max_{r<pool_size} in[b, (q*pool_step + r)*S + j]),  S = pool_stride
 * maxpool bwd: in_diff[b, p*S + j] = (sum_q [in == out_q] * out_diff_q) / #pools containing patch p */
int aslp_conv_gather_patches(aslp_stream_t s, float* patches, int ldp, const float* in, int ldi, int rows, int num_patches, int num_splice,
                             int patch_dim, int patch_step, int patch_stride);
int aslp_conv_scatter_patch_diffs(aslp_stream_t s, float* in_diff, int ldd, const float* patch_diffs, int ldp, int rows, int num_patches,
                                  int num_splice, int patch_dim, int patch_step, int patch_stride);
int aslp_maxpool_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int num_pools, int pool_size, int pool_step,
                     int pool_stride);
int aslp_maxpool_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff,
                     int ldod, int rows, int num_patches, int num_pools, int pool_size, int pool_step, int pool_stride);

/* ---- CompactFsmn memory block (src/aslp-nnet/nnet-cfsmn-component.h:169-264) ----
 * fwd : out[t,d] = in[t,d] + sum_{c=0}^{P+F} coef[c,d] * in[t+c-P, d]        (zero outside [0,T))
 * bwd : in_diff[t,d] = out_diff[t,d] + sum_c coef[P+F-c, d] * out_diff[t+c-F, d]
 * grad: coef_corr[c,d] = sum_t in[t+c-P,d] * out_diff[t,d]  (beta = 0, :224), then optional clip. */
int aslp_fsmn_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int T, int D,
                  const float* coef, int ldc, int past, int future);
int aslp_fsmn_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* out_diff, int ldo, int T, int D,
                  const float* coef, int ldc, int past, int future);
int aslp_fsmn_coef_grad(aslp_stream_t s, float* coef_corr, int ldc, const float* in, int ldi,
                        const float* out_diff, int ldo, int T, int D, int past, int future, float clip);

/* ---- the smaller components of the zoo (csrc/zoo.cu) ----
 * Dropout (src/aslp-nnet/nnet-activation.h:203-273): mask = (u < retention), out = in * mask / retention; the mask is kept (as
 * 0/1 floats, like the reference's dropout_mask_) for the backward pass.  u comes from Philox-4x32-10 keyed by `seed`, counter =
 * (element index / 4, call): the mask depends only on (seed, call, rows, cols), not on the launch geometry. */
int aslp_dropout_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, float* mask, int ldm, int rows, int cols,
                     float retention, unsigned long long seed, unsigned long long call);
/* out = a * b * scale (MulElements + Scale: Dropout with a host-drawn mask, Dropout backward) */
int aslp_mul_elements(aslp_stream_t s, float* out, int ldo, const float* a, int lda, const float* b, int ldb, int rows, int cols, float scale);
/* BlockSoftmax backward of one block (nnet-activation.h:120-139): dst = src * (1 - sum over the block's columns of src) */
int aslp_rows_one_minus_sum(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int cols);
/* PnormComponent / MaxoutComponent (nnet-activation.h:305-377; MatrixBase::GroupPnorm, GroupPnormDeriv, GroupMax, GroupMaxDeriv,
 * MulRowsGroupMat, kaldi-matrix.cc:1071-1138, 2530-2558): in [rows, groups * group_size] -> out [rows, groups] */
int aslp_group_pnorm_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int groups, int group_size, float p);
int aslp_group_pnorm_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff, int ldod,
                         int rows, int groups, int group_size, float p);
int aslp_group_max_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, int rows, int groups, int group_size);
int aslp_group_max_bwd(aslp_stream_t s, float* in_diff, int ldd, const float* in, int ldi, const float* out, int ldo, const float* out_diff, int ldod,
                       int rows, int groups, int group_size);
/* LengthNormComponent (nnet-various.h:327-365): row_scales[r] = 1 / |x_r|_2, out = x * row_scales; backward = aslp_mul_rows_vec */
int aslp_length_norm_fwd(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, float* row_scales, int rows, int cols);
int aslp_mul_rows_vec(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, const float* v, int rows, int cols);
/* CopyComponent (nnet-various.h:186-316, cu::Copy): out[r][c] = in[r][idx[c]] */
int aslp_copy_cols(aslp_stream_t s, float* out, int ldo, const float* in, int ldi, const int* idx_dev, int rows, int cols_out);
/* LstmCifgProjectedStreams (nnet-lstm-couple-if-projected-streams.h): rows [g | f | o] -> [g | -f | f | o] of a parameter matrix
 * (i = 1 - f = sigmoid(-pre_f), so the four-gate recurrence serves the coupled cell), and the derivative columns
 * [dg | di | df | do] -> [dg | df - di | do] for the three-gate weight gradients */
int aslp_cifg_expand(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int C, int cols);
int aslp_cifg_compact(aslp_stream_t s, float* dst, int ldd, const float* src, int lds, int rows, int C);

/* ---- LSTM family recurrence: one persistent launch walks all T steps ----
 * Lstm (nnet-recurrent-component.cc:235-480), LstmProjectedStreams (nnet-lstm-projected-streams.h:313-617),
 * BLstm / BLstmProjectedStreams (nnet-blstm-projected-streams.h:468-), BLstmProjectedStreamsLC
 * (nnet-blstm-projected-streams-lc.h:503-1082).  Buffers keep the reference layout:
 * [(T+2)*S, 7C+R] rows t*S+s, columns [g i f o c h m r]; row block 0 / T+1 are the boundary states.
 * Before fwd the caller has put x*W_x^T + bias into the gifo columns of rows [S,(T+1)S) (one GEMM).
 * Before bwd the caller has put out_diff into the r (or m if R==0) columns of dbuf rows [S,(T+1)S). */
typedef struct {
  int T, S, C, R;            /* R == 0: no projection, the recurrent input is m and w_r is [4C, C] */
  int reverse;               /* 0: t = 1..T reading t-1;  1: t = T..1 reading t+1 */
  float* buf;  int ldb;      /* forward activations  [(T+2)S, 7C+R] */
  float* dbuf; int lddb;     /* backward derivatives [(T+2)S, 7C+R] (bwd only; NULL for fwd) */
  const float* w_r;  int ldwr;   /* [4C, R or C] */
  const float* w_rm; int ldwrm;  /* [R, C], NULL if R == 0 */
  const float* peep_i; const float* peep_f; const float* peep_o; /* [C] each */
  const int* seq_len_dev;    /* int[S] or NULL: fwd zeroes row t of a stream when t > len (blstm-projected-streams.h:654-657) */
  float cell_clip;           /* 50 (ApplyFloor(-50)/ApplyCeiling(50)) */
} aslp_lstm_dir_t;
size_t aslp_lstm_workspace_bytes(int T, int S, int C, int R, int ndirs, int backward);
int aslp_lstm_seq_fwd(aslp_stream_t s, const aslp_lstm_dir_t* dirs, int ndirs, void* workspace, size_t workspace_bytes);
int aslp_lstm_seq_bwd(aslp_stream_t s, const aslp_lstm_dir_t* dirs, int ndirs, void* workspace, size_t workspace_bytes);
/* measurement aid: CUDA-event timing of the persistent launches on their own stream (enable, run, read totals in ms) */
int aslp_lstm_profile(int enable);
int aslp_lstm_profile_read(double* fwd_ms, int* fwd_launches, double* bwd_ms, int* bwd_launches);
/* development aid: when dev_buf != NULL (long long[grid][12], device) thread 0 of every CTA accumulates clock64 ticks spent
 * {waiting for its CTA, in its own exchange poll, waiting for the CTA's polls, in the work units, [fwd: contraction, reduce, unit prologue]} over the launch */
int aslp_lstm_debug_timing(long long* dev_buf);

/* ---- GruStreams recurrence (src/aslp-nnet/nnet-gru-streams.h:238-441) ----
 * buf [(T+2)S, 5H] columns [z r m g h]; before fwd rows [S,(T+1)S) hold x*W_zrm_x^T + bias in [z r m]. */
typedef struct {
  int T, S, H;
  float* buf;  int ldb;
  float* dbuf; int lddb;      /* bwd only: [(T+2)S, 5H] with out_diff preloaded into the h columns */
  const float* w_zr_h; int ldwzr;   /* [2H, H] */
  const float* w_m_g;  int ldwmg;   /* [H, H] */
} aslp_gru_t;
size_t aslp_gru_workspace_bytes(int T, int S, int H, int backward);
int aslp_gru_seq_fwd(aslp_stream_t s, const aslp_gru_t* g, void* workspace, size_t workspace_bytes);
int aslp_gru_seq_bwd(aslp_stream_t s, const aslp_gru_t* g, void* workspace, size_t workspace_bytes);

/* ---- Eesen-style CTC on probabilities (src/aslp-nnet/ctc-loss.cc:115-227; cu-kernels.cu:3276-3534) ----
 * probs [T*S, K] stream-interleaved softmax outputs; labels_dev int[S, 2*Lmax+1] expanded with blanks
 * and padded with -1 (ctc-loss.cc:133-150); frames per stream in seq_len_dev.
 * Outputs: pzx_dev[S] = log p(z|x); diff [T*S, K] = gradient w.r.t. the pre-softmax activations. */
size_t aslp_ctc_eesen_workspace_bytes(int T, int S, int K, int Lexp_max);
int aslp_ctc_eesen(aslp_stream_t s, float* diff, int ldd, const float* probs, int ldp, int T, int S, int K,
                   const int* labels_dev, int Lexp_max, const int* seq_len_dev, float* pzx_dev,
                   void* workspace, size_t workspace_bytes);

/* ---- aslp-parallel averaging kernels (src/aslp-parallel/{bsp,bmuf,sod}-worker.cc, optimizer.h) ----
 * All operate on ONE packed fp32 arena of n elements (UpdatableComponent::GetGpuParams order). */
/* BSP pre-scale: buf = w * factor (bsp-worker.cc:41-47) */
int aslp_sync_scale(aslp_stream_t s, float* dst, const float* w, size_t n, float factor);
/* BMUF: g = w - w_prev (bmuf-worker.cc:46-50) */
int aslp_sync_diff(aslp_stream_t s, float* g, const float* a, const float* b, size_t n);
/* BMUF filter after the sum-allreduce of g (bmuf-worker.cc:55-66):
 *   delta = mom*delta_prev + (1-mom)*lr*g ; w = w_prev + delta ; w_prev = w ; delta_prev = delta */
int aslp_sync_bmuf_apply(aslp_stream_t s, float* w, float* w_prev, float* delta_prev, const float* g_sum,
                         size_t n, float momentum, float learn_rate);
/* SOD optimizers applied to the summed difference g = w_prev - w (sod-worker.cc:46-60; optimizer.h:21-170) */
enum { ASLP_OPT_SGD = 0, ASLP_OPT_MOMENTUM = 1, ASLP_OPT_ADAGRAD = 2, ASLP_OPT_RMSPROP = 3, ASLP_OPT_ADADELTA = 4, ASLP_OPT_ADAM = 5 };
int aslp_sync_sod_apply(aslp_stream_t s, int opt, float* w, const float* g, float* state1, float* state2,
                        size_t n, float lr, float p1, float p2, float eps, int step);

/* multi-tensor forms: parameters stay in the components' own tensors (UpdatableComponent::GetGpuParams views, element
 * counts including row padding); `table_dev` is a DEVICE array of {ptr, offset into the packed arena, n}.  One launch each. */
typedef struct { float* ptr; size_t offset; size_t n; } aslp_tensor_ref_t;
int aslp_sync_pack(aslp_stream_t s, float* arena, const aslp_tensor_ref_t* table_dev, int ntensors, float factor);           /* arena = w * factor (BSP) */
/* arena = w * (float(frames) / float(*frames_all_dev)): the BSP weight of bsp-worker.cc:44 with the job's frame total still on the
 * device (behind its all-reduce on the same stream), so that a pipelined exchange needs no host round trip for it */
int aslp_sync_pack_weighted(aslp_stream_t s, float* arena, const aslp_tensor_ref_t* table_dev, int ntensors, int frames, const int* frames_all_dev);
int aslp_sync_pack_diff(aslp_stream_t s, float* arena, const aslp_tensor_ref_t* table_dev, int ntensors, const float* w_prev_arena, float sign);  /* arena = sign*(w - w_prev) */
int aslp_sync_unpack(aslp_stream_t s, const float* arena, const aslp_tensor_ref_t* table_dev, int ntensors);                 /* w = arena */
int aslp_sync_bmuf_apply_packed(aslp_stream_t s, const aslp_tensor_ref_t* table_dev, int ntensors, float* w_prev_arena, float* delta_prev_arena,
                                const float* g_sum_arena, float momentum, float learn_rate);
int aslp_sync_sod_apply_packed(aslp_stream_t s, int opt, const aslp_tensor_ref_t* table_dev, int ntensors, const float* g_sum_arena, float* state1,
                               float* state2, float* w_prev_arena, float lr, float p1, float p2, float eps, int step);

/* NCCL communicator owned by the library (replaces MpiNode, src/aslp-parallel/mpi-node.h:18-101) */
typedef struct aslp_comm* aslp_comm_t;
int aslp_comm_unique_id(char id_out[128]);                       /* rank 0, then shipped by the launcher */
int aslp_comm_init(aslp_comm_t* c, const char id[128], int nranks, int rank);
int aslp_comm_destroy(aslp_comm_t c);
int aslp_comm_rank(aslp_comm_t c, int* rank, int* nranks);
int aslp_comm_allreduce_sum_f32(aslp_comm_t c, aslp_stream_t s, float* buf, size_t n);   /* MpiNode::AllReduce (mpi-node.h:69-73) */
int aslp_comm_allreduce_sum_f64(aslp_comm_t c, aslp_stream_t s, double* buf, size_t n);  /* ReduceAccStat (mpi-node.h:76-91) */
int aslp_comm_allreduce_sum_i32(aslp_comm_t c, aslp_stream_t s, int* buf, size_t n);
int aslp_comm_barrier(aslp_comm_t c, aslp_stream_t s);
/* point-to-point exchange of a packed fp32 arena with one peer: the async parameter-server modes
 * (MPI_Send / MPI_Recv / MPI_Sendrecv in easgd-worker.cc:49-56, easgd-server.cc:66-73, asgd-worker.cc:47-58, asgd-server.cc:82-99) */
int aslp_comm_send_f32(aslp_comm_t c, aslp_stream_t s, const float* buf, size_t n, int peer);
int aslp_comm_recv_f32(aslp_comm_t c, aslp_stream_t s, float* buf, size_t n, int peer);
int aslp_comm_sendrecv_f32(aslp_comm_t c, aslp_stream_t s, const float* sendbuf, float* recvbuf, size_t n, int peer);

#ifdef __cplusplus
}
#endif
#endif /* ASLP_B200_H_ */
