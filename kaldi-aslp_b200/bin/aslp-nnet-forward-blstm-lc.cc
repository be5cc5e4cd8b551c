// aslp-nnet-forward-blstm-lc -- Feedforward of a latency-controlled BLSTM net, chunk by chunk with the forward state
// carried across chunks, as src/aslp-nnetbin/aslp-nnet-forward-blstm-lc.cc:30-243: same flags, chunk arithmetic and log
// lines.  Two quirks of the reference are kept on purpose (its output depends on both):
//  * batch_size = chunk_size + right_splice is evaluated BEFORE the command line is parsed (:50-52 vs :73), i.e. it is
//    always 64 + 16 = 80 rows whatever --chunk-size / --right-splice say; the flags only move the chunk offsets, the copy
//    length and the row the forward state is carried from.  With short utterances the backward-direction chain therefore
//    sees every remaining frame, not right_splice of them (found by matching the reference archives to 1e-7).
//  * the [80 x dim] input buffer is zeroed once per utterance, so rows past the frames of a chunk still hold what earlier
//    chunks put there and feed the backward-direction chain (:160-170).
#include "nnet-nnet.h"
#include "nnet-pdf-prior.h"
#include "parse-options.h"
#include "table.h"

int main(int argc, char* argv[]) {
  using namespace kaldi;
  using namespace kaldi::aslp_nnet;
  try {
    const char* usage =
        "Perform forward pass for Latency Control BLSTM through Neural Network.\n"
        "\n"
        "Usage:  aslp-nnet-forward-blstm-lc [options] <model-in> <feature-rspecifier> <feature-wspecifier>\n"
        "e.g.: \n"
        " aslp-nnet-forward-blstm-lc nnet ark:features.ark ark:mlpoutput.ark\n";
    ParseOptions po(usage);
    PdfPriorOptions prior_opts;
    prior_opts.Register(&po);
    int32 chunk_size = 64;
    po.Register("chunk-size", &chunk_size, "---BLSTM--- Latency-controlled BPTT chunk size, must be same with training");
    int32 right_splice = 16;
    po.Register("right-splice", &right_splice, "---BLSTM--- Latency-controlled BPTT right context size, must be same with training");
    std::string feature_transform;
    po.Register("feature-transform", &feature_transform, "Feature transform in front of main network (in nnet format)");
    bool no_softmax = false;
    po.Register("no-softmax", &no_softmax, "No softmax on MLP output (or remove it if found), the pre-softmax activations will be used as log-likelihoods, log-priors will be subtracted");
    bool apply_log = true;
    po.Register("apply-log", &apply_log, "Transform MLP output to logscale");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    int32 gpu_id = -1;
    po.Register("gpu-id", &gpu_id, "selected gpu id, if negative then select automaticly");
    const int32 batch_size = chunk_size + right_splice;      // quirk: the DEFAULTS (64 + 16), taken before po.Read()
    po.Read(argc, argv);
    if (po.NumArgs() != 3) { po.PrintUsage(); return 1; }
    const std::string model_filename = po.GetArg(1), feature_rspecifier = po.GetArg(2), feature_wspecifier = po.GetArg(3);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (gpu_id >= 0) ASLP_OK(aslp_set_device(gpu_id));

    Nnet nnet_transf;
    if (feature_transform != "") nnet_transf.Read(feature_transform);
    Nnet nnet;
    nnet.Read(model_filename);
    if (apply_log && no_softmax) KALDI_ERR << "Cannot use both --apply-log=true --no-softmax=true, use only one of the two!";
    PdfPrior pdf_prior(prior_opts);
    nnet.SetChunkSize(chunk_size);

    int64 tot_t = 0;
    SequentialBaseFloatMatrixReader feature_reader(feature_rspecifier);
    BaseFloatMatrixWriter feature_writer(feature_wspecifier);
    CuMatrix<BaseFloat> feats, feats_transf, nnet_in, nnet_out, nnet_out_chunk;
    Matrix<BaseFloat> nnet_out_host;
    const int32 feat_dim = nnet.InputDim(), out_dim = nnet.OutputDim();
    Timer time;
    int32 num_done = 0;
    for (; !feature_reader.Done(); feature_reader.Next()) {
      const Matrix<BaseFloat>& mat = feature_reader.Value();
      const std::string utt = feature_reader.Key();
      KALDI_VLOG(2) << "Processing utterance " << num_done + 1 << ", " << utt << ", " << mat.NumRows() << "frm";
      double sum = 0.0;
      for (int32 r = 0; r < mat.NumRows(); r++) for (int32 c = 0; c < mat.NumCols(); c++) sum += mat.RowData(r)[c];
      if (!KALDI_ISFINITE(sum)) KALDI_ERR << "NaN or inf found in features for " << utt;
      feats = mat;
      const CuMatrixBase<BaseFloat>* net_in = &feats;
      if (nnet_transf.NumComponents() > 0) {
        nnet_transf.Feedforward(feats, &feats_transf);
        if (!KALDI_ISFINITE(feats_transf.Sum())) KALDI_ERR << "NaN or inf found in transformed-features for " << utt;
        net_in = &feats_transf;
      }
      // new utterance: history state reset
      std::vector<int32> reset_flags(1, 1);
      nnet.ResetLstmStreams(reset_flags);
      const int32 num_frames = net_in->NumRows();
      const int32 num_chunks = (num_frames - 1) / chunk_size + 1;
      nnet_out.Resize(num_frames, out_dim);
      nnet_in.Resize(batch_size, feat_dim);          // zeroed once per utterance, NOT per chunk (quirk, see the header)
      for (int32 i = 0; i < num_chunks; i++) {
        const int32 offset = i * chunk_size;
        const int32 len = offset + batch_size < num_frames ? batch_size : num_frames - offset;
        const int32 copy_len = offset + chunk_size < num_frames ? chunk_size : num_frames - offset;
        KALDI_ASSERT(len <= batch_size);
        nnet_in.RowRange(0, len).CopyFromMat(net_in->RowRange(offset, len));
        nnet.Feedforward(nnet_in, &nnet_out_chunk);
        nnet_out.RowRange(offset, copy_len).CopyFromMat(nnet_out_chunk.RowRange(0, copy_len));
      }
      if (!KALDI_ISFINITE(nnet_out.Sum())) KALDI_ERR << "NaN or inf found in nn-output for " << utt;
      FinalizePosteriors(utt, apply_log, 0.0f, prior_opts.class_frame_counts, pdf_prior, &nnet_out);
      nnet_out.CopyToMat(&nnet_out_host);
      feature_writer.Write(utt, nnet_out_host);
      if (num_done % 100 == 0) {
        const double time_now = time.Elapsed();
        KALDI_VLOG(1) << "After " << num_done << " utterances: time elapsed = " << time_now / 60 << " min; processed "
                      << tot_t / time_now << " frames per second.";
      }
      num_done++;
      tot_t += mat.NumRows();
    }
    KALDI_LOG << "Done " << num_done << "files" << " in " << time.Elapsed() / 60 << "min," << " (fps " << tot_t / time.Elapsed() << ")";
    if (num_done == 0) return -1;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}
