// nnet-pdf-prior.h -- PdfPriorOptions / PdfPrior with the reference's flags and arithmetic
// (src/aslp-nnet/nnet-pdf-prior.h:35-74, nnet-pdf-prior.cc:27-86): log-priors from a text vector of class frame counts
// (double arithmetic on the host, classes under --prior-floor get sqrt(FLT_MAX)), kept on the device; the subtraction itself
// is fused into aslp_posterior_finalize (one pass with the log / blank stages of the forwarders).
#ifndef ASLP_HOST_NNET_PDF_PRIOR_H_
#define ASLP_HOST_NNET_PDF_PRIOR_H_
#include <cfloat>
#include <cmath>
#include "io.h"
#include "matrix.h"
#include "parse-options.h"

namespace kaldi {
namespace aslp_nnet {

struct PdfPriorOptions {
  std::string class_frame_counts;
  BaseFloat prior_scale;
  BaseFloat prior_floor;
  PdfPriorOptions() : class_frame_counts(""), prior_scale(1.0), prior_floor(1e-10) {}
  void Register(ParseOptions* opts) {
    opts->Register("class-frame-counts", &class_frame_counts, "Vector with frame-counts of pdfs to compute log-priors."
                   " (priors are typically subtracted from log-posteriors or pre-softmax activations)");
    opts->Register("prior-scale", &prior_scale, "Scaling factor to be applied on pdf-log-priors");
    opts->Register("prior-floor", &prior_floor, "Flooring constatnt for prior probability (i.e. label rel. frequency)");
  }
};

class PdfPrior {
 public:
  explicit PdfPrior(const PdfPriorOptions& opts) : prior_scale_(opts.prior_scale) {
    if (opts.class_frame_counts == "") return;          // deactivated (e.g. bottleneck features)
    KALDI_LOG << "Computing pdf-priors from : " << opts.class_frame_counts;
    Vector<double> frame_counts;
    {
      Input in;
      in.OpenTextMode(opts.class_frame_counts);
      frame_counts.Read(in.Stream(), false);
      in.Close();
    }
    double sum = 0.0;
    for (int32 i = 0; i < frame_counts.Dim(); i++) sum += frame_counts(i);
    Vector<BaseFloat> log_priors(frame_counts.Dim());
    int32 num_floored = 0;
    double check = 0.0;
    for (int32 i = 0; i < frame_counts.Dim(); i++) {
      const double rel_freq = frame_counts(i) * (1.0 / sum);
      double lp = std::log(rel_freq + 1e-20);
      if (rel_freq < opts.prior_floor) { lp = std::sqrt(FLT_MAX); num_floored++; }
      check += lp;
      log_priors(i) = static_cast<BaseFloat>(lp);
    }
    KALDI_LOG << "Floored " << num_floored << " pdf-priors (hard-set to " << std::sqrt(FLT_MAX)
              << ", which disables DNN output when decoding)";
    KALDI_ASSERT(std::isfinite(check));
    log_priors_ = log_priors;
  }
  int32 Dim() const { return log_priors_.Dim(); }
  BaseFloat PriorScale() const { return prior_scale_; }
  // nnet-pdf-prior.cc:75-86, for code written against the reference class: llk -= prior_scale * log_priors on every row
  void SubtractOnLogpost(CuMatrixBase<BaseFloat>* llk) {
    const float* lp = DeviceLogPriors(llk->NumCols());
    ASLP_OK(aslp_add_vec_to_rows(CuStream(), llk->Data(), llk->Stride(), llk->NumRows(), llk->NumCols(), lp, -prior_scale_, 1.0f));
  }
  // device pointer for aslp_posterior_finalize; same two errors as PdfPrior::SubtractOnLogpost (:73-84)
  const float* DeviceLogPriors(int32 num_cols) const {
    if (log_priors_.Dim() == 0) KALDI_ERR << "--class-frame-counts is empty: Cannot initialize priors without the counts.";
    if (log_priors_.Dim() != num_cols)
      KALDI_ERR << "Dimensionality mismatch, class_frame_counts " << log_priors_.Dim() << " pdf_output_llk " << num_cols;
    return log_priors_.Data();
  }
 private:
  BaseFloat prior_scale_;
  CuVector<BaseFloat> log_priors_;
};

// The tail of the forwarders (aslp-nnet-forward.cc:184-216, -forward-blstm-lc.cc:178-199): log, blank scaling, prior
// subtraction, the two "doesn't look like probabilities" warnings and the finiteness check, in one device pass.
inline void FinalizePosteriors(const std::string& utt, bool apply_log, BaseFloat scale_blank, const std::string& class_frame_counts,
                               const PdfPrior& pdf_prior, CuMatrix<BaseFloat>* nnet_out) {
  static float* stats_dev = nullptr;
  if (stats_dev == nullptr) ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&stats_dev), 8 * sizeof(float)));
  const bool use_prior = class_frame_counts != "";
  const float* lp = use_prior ? pdf_prior.DeviceLogPriors(nnet_out->NumCols()) : nullptr;
  ASLP_OK(aslp_posterior_finalize(CuStream(), nnet_out->Data(), nnet_out->Stride(), nnet_out->NumRows(), nnet_out->NumCols(),
                                  apply_log ? 1 : 0, 1e-20f, scale_blank, lp, pdf_prior.PriorScale(), stats_dev));
  float h[5];
  ASLP_OK(aslp_memcpy_d2h(CuStream(), h, stats_dev, sizeof(h)));
  CuSync();
  if (apply_log && !(h[0] >= 0.0f && h[1] <= 1.0f))
    KALDI_WARN << utt << " Applying 'log' to data which don't seem to be probabilities (is there a softmax somwhere?)";
  if (use_prior && h[2] >= 0.0f && h[3] <= 1.0f)
    KALDI_WARN << utt << " Subtracting log-prior on 'probability-like' data in range [0..1] "
               << "(Did you forget --no-softmax=true or --apply-log=true ?)";
  if (h[4] > 0.0f) KALDI_ERR << "NaN or inf found in final output nn-output for " << utt;
}

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
