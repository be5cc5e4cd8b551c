// forward-main.h -- the body shared by aslp-nnet-forward and aslp-nnet-forward-skip: Feedforward pass of a trained net over a
// feature table (cross-validation / decoding front end), same command line, option checks, post-processing order and log
// lines as src/aslp-nnetbin/aslp-nnet-forward.cc:30-266 and aslp-nnet-forward-skip.cc:33-251.  The two reference mains differ
// only in what --skip-width does: `forward` runs the net on every skip_width-th frame and repeats each output row
// skip_width times (:161-177); `forward-skip` runs it once per offset 0..skip_width-1 on the frames offset + i*skip_width
// and scatters the outputs back to those rows (forward-skip.cc:150-171), so every frame gets its own posterior.
// The post-processing (log, blank scaling, prior subtraction, range warnings, finiteness check) is ONE device pass
// (aslp_posterior_finalize) instead of nine matrix methods; the strided row gather / scatter of the skip forms is one 2-D
// copy each instead of a per-row loop.  --use-gpu=no is refused: this build has no CPU path.
#ifndef ASLP_BIN_FORWARD_MAIN_H_
#define ASLP_BIN_FORWARD_MAIN_H_
#include "nnet-nnet.h"
#include "nnet-pdf-prior.h"
#include "parse-options.h"
#include "table.h"

namespace kaldi {
namespace aslp_nnet {

inline int ForwardMain(int argc, char* argv[], bool split_skip) {
  try {
    const std::string name = split_skip ? "aslp-nnet-forward-skip" : "aslp-nnet-forward";
    const std::string usage =
        "Perform forward pass through Neural Network.\n"
        "\n"
        "Usage:  " + name + " [options] <model-in> <feature-rspecifier> <feature-wspecifier>\n"
        "e.g.: \n"
        " " + name + " nnet ark:features.ark ark:mlpoutput.ark\n";
    ParseOptions po(usage.c_str());
    PdfPriorOptions prior_opts;
    prior_opts.Register(&po);
    std::string feature_transform;
    po.Register("feature-transform", &feature_transform, "Feature transform in front of main network (in nnet format)");
    bool no_softmax = false;
    po.Register("no-softmax", &no_softmax, "No softmax on MLP output (or remove it if found), the pre-softmax activations will be used as log-likelihoods, log-priors will be subtracted");
    bool apply_log = true;
    po.Register("apply-log", &apply_log, "Transform MLP output to logscale");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    bool add_softmax = false;
    po.Register("add-softmax", &add_softmax, "add softmax calulation for warp-ctc training");
    int32 time_shift = 0;
    po.Register("time-shift", &time_shift, "LSTM : repeat last input frame N-times, discrad N initial output frames.");
    float scale_blank = 0.0;
    po.Register("scale-blank", &scale_blank, "scale the blank posterior for CTC decoding");
    int32 skip_width = 0;
    po.Register("skip-width", &skip_width, "num of frame for one skip(default 0, not use skip)");
    int32 gpu_id = -1;
    po.Register("gpu-id", &gpu_id, "selected gpu id, if negative then select automaticly");
    po.Read(argc, argv);
    if (po.NumArgs() != 3) { po.PrintUsage(); return 1; }
    const std::string model_filename = po.GetArg(1), feature_rspecifier = po.GetArg(2), feature_wspecifier = po.GetArg(3);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (gpu_id >= 0) ASLP_OK(aslp_set_device(gpu_id));

    Nnet nnet_transf;
    if (feature_transform != "") nnet_transf.Read(feature_transform);
    Nnet nnet;
    nnet.Read(model_filename);
    // avoid some bad option combinations (:103-105; the reference's softmax removal is commented out there, so
    // --no-softmax only takes part in this check)
    if (apply_log && no_softmax) KALDI_ERR << "Cannot use both --apply-log=true --no-softmax=true, use only one of the two!";
    PdfPrior pdf_prior(prior_opts);

    int64 tot_t = 0;
    SequentialBaseFloatMatrixReader feature_reader(feature_rspecifier);
    BaseFloatMatrixWriter feature_writer(feature_wspecifier);
    CuMatrix<BaseFloat> feats, feats_transf, nnet_out, skip_feat, skip_out, tmp_out;
    Matrix<BaseFloat> nnet_out_host;
    Timer time;
    int32 num_done = 0;
    for (; !feature_reader.Done(); feature_reader.Next()) {
      const Matrix<BaseFloat>& in = feature_reader.Value();
      const std::string utt = feature_reader.Key();
      KALDI_VLOG(2) << "Processing utterance " << num_done + 1 << ", " << utt << ", " << in.NumRows() << "frm";
      // time-shift: repeat the last input frame N times (:139-145)
      Matrix<BaseFloat> mat(in.NumRows() + (time_shift > 0 ? time_shift : 0), in.NumCols());
      double sum = 0.0;
      for (int32 r = 0; r < mat.NumRows(); r++) {
        const float* src = in.RowData(std::min(r, in.NumRows() - 1));
        std::copy(src, src + in.NumCols(), mat.RowData(r));
        if (r < in.NumRows()) for (int32 c = 0; c < in.NumCols(); c++) sum += src[c];
      }
      if (!KALDI_ISFINITE(sum)) KALDI_ERR << "NaN or inf found in features for " << utt;
      feats = mat;
      const CuMatrixBase<BaseFloat>* net_in = &feats;
      if (nnet_transf.NumComponents() > 0) {
        nnet_transf.Feedforward(feats, &feats_transf);
        if (!KALDI_ISFINITE(feats_transf.Sum())) KALDI_ERR << "NaN or inf found in transformed-features for " << utt;
        net_in = &feats_transf;
      }
      std::vector<int32> frame_num_utt;
      auto strided_rows = [&](CuMatrix<BaseFloat>* dst, int32 dst_row0, int32 dst_step, const CuMatrixBase<BaseFloat>& src, int32 src_row0, int32 src_step, int32 n) {
        // n rows: dst[dst_row0 + i*dst_step] = src[src_row0 + i*src_step] in ONE strided device copy
        ASLP_OK(aslp_memcpy2d_d2d(CuStream(), dst->Data() + static_cast<size_t>(dst_row0) * dst->Stride(), sizeof(float) * dst->Stride() * dst_step,
                                  src.Data() + static_cast<size_t>(src_row0) * src.Stride(), sizeof(float) * src.Stride() * src_step,
                                  sizeof(float) * src.NumCols(), n));
      };
      if (split_skip) {
        // split skip prediction (forward-skip.cc:150-171); skip_width <= 0 leaves no output at all there (the matrix methods
        // that follow assert on the empty matrix): refuse it by name
        if (skip_width < 1) KALDI_ERR << "--skip-width must be at least 1 for " << name;
        for (int32 skip_offset = 0; skip_offset < skip_width; skip_offset++) {
          const int32 skip_len = (net_in->NumRows() - 1 - skip_offset) / skip_width + 1;
          if (net_in->NumRows() - 1 - skip_offset < 0) break;        // fewer frames than offsets (the reference would resize to 0 rows and assert)
          skip_feat.Resize(skip_len, net_in->NumCols(), kUndefined);
          strided_rows(&skip_feat, 0, 1, *net_in, skip_offset, skip_width, skip_len);
          frame_num_utt.assign(1, skip_feat.NumRows());
          nnet.SetSeqLengths(frame_num_utt);
          nnet.Feedforward(skip_feat, &skip_out);
          if (nnet_out.NumRows() != net_in->NumRows() || nnet_out.NumCols() != skip_out.NumCols())
            nnet_out.Resize(net_in->NumRows(), skip_out.NumCols(), kSetZero);
          strided_rows(&nnet_out, skip_offset, skip_width, skip_out, 0, 1, skip_len);
        }
      } else if (skip_width > 1) {
        // skip prediction (:161-177): every skip_width-th frame goes through the net, outputs are repeated
        const int32 skip_len = (net_in->NumRows() - 1) / skip_width + 1;
        skip_feat.Resize(skip_len, net_in->NumCols(), kUndefined);
        strided_rows(&skip_feat, 0, 1, *net_in, 0, skip_width, skip_len);
        frame_num_utt.push_back(skip_feat.NumRows());
        nnet.SetSeqLengths(frame_num_utt);
        nnet.Feedforward(skip_feat, &skip_out);
        nnet_out.Resize(net_in->NumRows(), skip_out.NumCols(), kUndefined);
        for (int32 j = 0; j < skip_width; j++) {
          const int32 n = (nnet_out.NumRows() - 1 - j) / skip_width + 1;
          if (nnet_out.NumRows() - 1 - j >= 0) strided_rows(&nnet_out, j, skip_width, skip_out, 0, 1, n);
        }
      } else {
        frame_num_utt.push_back(net_in->NumRows());
        nnet.SetSeqLengths(frame_num_utt);
        nnet.Feedforward(*net_in, &nnet_out);
      }
      if (add_softmax) {                       // extra softmax for warp-ctc nets (:181-184)
        tmp_out = nnet_out;
        ASLP_OK(aslp_softmax_rows(CuStream(), nnet_out.Data(), nnet_out.Stride(), tmp_out.Data(), tmp_out.Stride(),
                                  nnet_out.NumRows(), nnet_out.NumCols()));
      }
      if (!KALDI_ISFINITE(nnet_out.Sum())) KALDI_ERR << "NaN or inf found in nn-output for " << utt;
      FinalizePosteriors(utt, apply_log, scale_blank, prior_opts.class_frame_counts, pdf_prior, &nnet_out);
      nnet_out.CopyToMat(&nnet_out_host);
      if (time_shift > 0) {                    // drop the N first output frames (:222-225)
        Matrix<BaseFloat> tmp(nnet_out_host.NumRows() - time_shift, nnet_out_host.NumCols());
        for (int32 r = 0; r < tmp.NumRows(); r++)
          std::copy(nnet_out_host.RowData(r + time_shift), nnet_out_host.RowData(r + time_shift) + tmp.NumCols(), tmp.RowData(r));
        nnet_out_host = tmp;
      }
      feature_writer.Write(utt, nnet_out_host);
      if (num_done % 100 == 0) {
        const double time_now = time.Elapsed();
        KALDI_VLOG(1) << "After " << num_done << " utterances: time elapsed = " << time_now / 60 << " min; processed "
                      << tot_t / time_now << " frames per second.";
      }
      num_done++;
      tot_t += mat.NumRows();
    }
    KALDI_LOG << "Done " << num_done << " files" << " in " << time.Elapsed() / 60 << "min," << " (fps " << tot_t / time.Elapsed() << ")";
    if (num_done == 0) return -1;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
