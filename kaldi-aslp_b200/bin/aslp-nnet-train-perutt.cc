// aslp-nnet-train-perutt -- one utterance per update (the FSMN / whole-sentence trainer, BASELINE config 4), same command
// line, length-tolerance rule, learn-rate quirk (learn_rate = norm_lr / 1024 for every utterance, :201) and log lines as
// src/aslp-nnetbin/aslp-nnet-train-perutt.cc:30-300.  --frame-weights is not served (not on the BASELINE configs).
#include <memory>
#include <algorithm>
#include "batch-feeder.h"
#include "nnet-nnet.h"
#include "nnet-train-step.h"
#include "nnet-loss.h"
#include "nnet-randomizer.h"
#include "nnet-trnopts.h"
#include "parse-options.h"
#include "table.h"

int main(int argc, char* argv[]) {
  using namespace kaldi;
  using namespace kaldi::aslp_nnet;
  try {
    const char* usage =
        "Perform one iteration of Neural Network training by Stochastic Gradient Descent.\n"
        "This version use pdf-posterior as targets, prepared typically by ali-to-post.\n"
        "The updates are done per-utterance, shuffling options are dummy for compatibility reason.\n"
        "\n"
        "Usage:  aslp-nnet-train-perutt [options] <feature-rspecifier> <targets-rspecifier> <model-in> [<model-out>]\n"
        "e.g.: \n"
        " aslp-nnet-train-perutt scp:feature.scp ark:posterior.ark nnet.init nnet.iter1\n";
    ParseOptions po(usage);
    NnetTrainOptions trn_opts;
    trn_opts.Register(&po);
    bool binary = true, crossvalidate = false;
    po.Register("binary", &binary, "Write output in binary mode");
    po.Register("cross-validate", &crossvalidate, "Perform cross-validation (don't backpropagate)");
    std::string feature_transform;
    po.Register("feature-transform", &feature_transform, "Feature transform in Nnet format");
    std::string objective_function = "xent";
    po.Register("objective-function", &objective_function, "Objective function : xent|mse");
    int32 length_tolerance = 5;
    po.Register("length-tolerance", &length_tolerance, "Allowed length difference of features/targets (frames)");
    std::string frame_weights;
    po.Register("frame-weights", &frame_weights, "Per-frame weights to scale gradients (frame selection/weighting).");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    NnetDataRandomizerOptions rnd_opts;      // dummy, for compatibility with the standard scripts
    rnd_opts.Register(&po);
    bool randomize = false;
    po.Register("randomize", &randomize, "Dummy option, for compatibility...");
    int32 report_period = 60000;
    po.Register("report-period", &report_period, "Number of frames for one report log, default(60000)");
    int32 drop_len = -1;
    po.Register("drop-len", &drop_len, "if sentence frame length greater than drop_len,if negative no drop");
    int32 gpu_id = -1;
    po.Register("gpu-id", &gpu_id, "selected gpu id, if negative then select automaticly");
    po.Read(argc, argv);
    if (po.NumArgs() != 4 - (crossvalidate ? 1 : 0)) { po.PrintUsage(); return 1; }
    const std::string feature_rspecifier = po.GetArg(1), targets_rspecifier = po.GetArg(2), model_filename = po.GetArg(3);
    std::string target_model_filename;
    if (!crossvalidate) target_model_filename = po.GetArg(4);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (gpu_id >= 0) ASLP_OK(aslp_set_device(gpu_id));
    if (frame_weights != "") KALDI_ERR << "--frame-weights is not supported by this build";

    Nnet nnet_transf;
    if (feature_transform != "") nnet_transf.Read(feature_transform);
    Nnet nnet;
    nnet.Read(model_filename);
    nnet.SetTrainOptions(trn_opts);
    const float norm_lr = trn_opts.learn_rate;
    long long total_frames = 0, report_frames = 0;
    SequentialBaseFloatMatrixReader feature_reader(feature_rspecifier);
    RandomAccessPosteriorReader targets_reader(targets_rspecifier);
    std::unique_ptr<LossItf> loss_holder;                  // LossItf* as in the reference's worker mains (xent | mse)
    if (objective_function == "xent") loss_holder.reset(new Xent);
    else if (objective_function == "mse") loss_holder.reset(new Mse);
    else KALDI_ERR << "Unsupported objective function: " << objective_function;
    LossItf& xent = *loss_holder;
    CuMatrix<BaseFloat> feats, feats_transf, nnet_out, obj_diff;
    Timer time;
    KALDI_LOG << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << " STARTED";
    int32 num_done = 0, num_no_tgt_mat = 0, num_other_error = 0;
    // One utterance per step, read / filtered / trimmed by the feeder thread as the reference's loop does it (:139-190)
    struct UttBatch {
      PinnedMatrix mat;
      Posterior targets;
      Vector<BaseFloat> weights;
    };
    auto fill = [&](UttBatch* b) -> bool {
      for (; !feature_reader.Done(); feature_reader.Next()) {
        const std::string utt = feature_reader.Key();
        if (!targets_reader.HasKey(utt)) {
          KALDI_WARN << utt << ", missing targets";
          num_no_tgt_mat++;                         // the feeder thread's until feeder.Join()
          continue;
        }
        const Matrix<BaseFloat>& full = feature_reader.Value();
        b->targets = targets_reader.Value(utt);
        // correct small length mismatch (:155-172)
        const int32 lens[3] = {full.NumRows(), static_cast<int32>(b->targets.size()), full.NumRows()};
        const int32 mn = *std::min_element(lens, lens + 3), mx = *std::max_element(lens, lens + 3);
        if (mx - mn >= length_tolerance) {
          KALDI_WARN << utt << ", length mismatch of targets " << b->targets.size() << " and features " << full.NumRows();
          num_other_error++;
          continue;
        }
        if (drop_len > 0 && mn > drop_len) {
          KALDI_WARN << utt << ", length too long " << mn << " drop it";
          continue;
        }
        b->mat.Resize(mn, full.NumCols(), kUndefined);
        for (int32 r = 0; r < mn; r++) std::copy(full.RowData(r), full.RowData(r) + full.NumCols(), b->mat.RowData(r));
        b->targets.resize(mn);
        b->weights.Resize(mn);
        for (int32 r = 0; r < mn; r++) b->weights(r) = 1.0f;
        feature_reader.Next();
        return true;
      }
      return false;
    };
    XentTrainStep train_step;
    Xent* xent_impl = dynamic_cast<Xent*>(loss_holder.get());
    BatchFeeder<UttBatch> feeder(fill, /*attach_device=*/false);
    while (UttBatch* b = feeder.Next()) {
      feats.Resize(b->mat.NumRows(), b->mat.NumCols(), kUndefined);
      feats.CopyFromHost(b->mat.Data(), b->mat.Stride());      // asynchronous: the slot is page-locked
      const CuMatrixBase<BaseFloat>* net_in = &feats;
      if (nnet_transf.NumComponents() > 0) { nnet_transf.Feedforward(feats, &feats_transf); net_in = &feats_transf; }
      trn_opts.learn_rate = norm_lr / 1024.0;             // quirk (:201): a fixed divisor, not the utterance length
      nnet.SetTrainOptions(trn_opts);
      if (!crossvalidate && xent_impl != nullptr) {
        // Propagate + Xent::Eval + Backpropagate; a step shape that repeats is recorded once and replayed (nnet-train-step.h)
        train_step.Run(&nnet, xent_impl, *net_in, b->weights, b->targets);
        feeder.Release(b);                                // weights and targets are staged
      } else {
        if (!crossvalidate) nnet.Propagate(*net_in, &nnet_out);
        else nnet.Feedforward(*net_in, &nnet_out);
        xent.Eval(b->weights, nnet_out, b->targets, &obj_diff);
        feeder.Release(b);                                // the loss has uploaded weights and targets
        if (!crossvalidate) nnet.Backpropagate(obj_diff, nullptr);
      }
      num_done++;
      total_frames += net_in->NumRows();
      report_frames += net_in->NumRows();
      if (report_frames >= report_period && report_period > 0) {
        KALDI_LOG << xent.Report();
        report_frames -= report_period;
      }
    }
    feeder.Join();
    if (!crossvalidate) nnet.Write(target_model_filename, binary);
    KALDI_LOG << "Done " << num_done << " files, " << num_no_tgt_mat << " with no tgt_mats, " << num_other_error << " with other errors. "
              << "[" << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << ", " << (randomize ? "RANDOMIZED" : "NOT-RANDOMIZED") << ", "
              << time.Elapsed() / 60 << " min, fps" << total_frames / time.Elapsed() << "]";
    KALDI_LOG << xent.Report();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}
