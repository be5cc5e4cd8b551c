// aslp-nnet-train-frame -- frame-shuffled minibatch training, same command line, bookkeeping and log lines as
// src/aslp-nnetbin/aslp-nnet-train-frame.cc:22-153.  --use-gpu=no is refused: this build has no CPU path.
// With --worker-type it is the worker of src/aslp-parallelbin/aslp-nnet-train-frame-worker.cc (bin/worker-opts.h).
#include <memory>
#include "nnet-nnet.h"
#include "nnet-train-step.h"
#include "nnet-loss.h"
#include "nnet-randomizer.h"
#include "nnet-trnopts.h"
#include "parse-options.h"
#include "table.h"
#include "worker-opts.h"

int main(int argc, char* argv[]) {
  using namespace kaldi;
  using namespace kaldi::aslp_nnet;
  try {
    const char* usage =
        "Perform one iteration of Neural Network training by mini-batch Stochastic Gradient Descent.\n"
        "Usage:  aslp-nnet-train-frame [options] <feature-rspecifier> <targets-rspecifier> <model-in> [<model-out>]\n"
        "e.g.: \n"
        " aslp-nnet-train-frame scp:feature.scp ark:posterior.ark nnet.init nnet.iter1\n";
    ParseOptions po(usage);
    NnetTrainOptions trn_opts;
    trn_opts.Register(&po);
    NnetDataRandomizerOptions rnd_opts;
    rnd_opts.Register(&po);
    bool binary = true, crossvalidate = false, randomize = true;
    po.Register("binary", &binary, "Write output in binary mode");
    po.Register("cross-validate", &crossvalidate, "Perform cross-validation (don't backpropagate)");
    po.Register("randomize", &randomize, "Perform the frame-level shuffling within the Cache::");
    std::string objective_function = "xent";
    po.Register("objective-function", &objective_function, "Objective function : xent|mse");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    double dropout_retention = 0.0;
    po.Register("dropout-retention", &dropout_retention, "number between 0..1, saying how many neurons to preserve (0.0 will keep original value");
    int32 report_period = -1;
    po.Register("report-period", &report_period, "Number of frames for one report log, default(-1, no report)");
    int32 gpu_id = -1;
    po.Register("gpu-id", &gpu_id, "selected gpu id, if negative then select automaticly");
    WorkerOptions wopts;
    wopts.Register(&po);
    po.Read(argc, argv);
    if (po.NumArgs() != 4 - (crossvalidate ? 1 : 0)) { po.PrintUsage(); return 1; }
    const std::string feature_rspecifier = po.GetArg(1), targets_rspecifier = po.GetArg(2), model_filename = po.GetArg(3);
    std::string target_model_filename;
    if (!crossvalidate) target_model_filename = po.GetArg(4);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (gpu_id >= 0) ASLP_OK(aslp_set_device(gpu_id));
    else wopts.SelectDevice();

    Nnet nnet;
    nnet.Read(model_filename);
    nnet.SetTrainOptions(trn_opts);
    // aslp-nnet-train-frame.cc:82-88: dropout retention from the command line while training, switched off (1.0) for cross-validation
    if (dropout_retention > 0.0) nnet.SetDropoutRetention(dropout_retention);
    if (crossvalidate) nnet.SetDropoutRetention(1.0);
    wopts.Create(&nnet, crossvalidate);
    std::unique_ptr<LossItf> loss_holder;                  // aslp-nnet-train-frame.cc:90-98
    if (objective_function == "xent") loss_holder.reset(new Xent);
    else if (objective_function == "mse") loss_holder.reset(new Mse);
    else KALDI_ERR << "Unsupported objective function: " << objective_function;
    LossItf& loss = *loss_holder;
    Timer time;
    long long total_frames = 0, report_frames = 0;
    KALDI_LOG << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << " STARTED";
    FrameDataReader reader(feature_rspecifier, targets_rspecifier, rnd_opts);
    const CuMatrixBase<BaseFloat>* nnet_in = nullptr;
    const Posterior* nnet_tgt = nullptr;
    CuMatrix<BaseFloat> nnet_out, obj_diff;
    XentTrainStep train_step;                              // the same three calls; a repeating minibatch shape is recorded once and replayed
    Xent* xent = dynamic_cast<Xent*>(loss_holder.get());
    Vector<BaseFloat> ones;
    while (!reader.Done()) {
      if (!reader.ReadData(&nnet_in, &nnet_tgt)) continue;
      if (!crossvalidate) wopts.BeforeBackpropagate(nnet_in->NumRows());      // a synchronisation due after this minibatch rides under it
      if (!crossvalidate && xent != nullptr) {
        if (ones.Dim() != nnet_in->NumRows()) { ones.Resize(nnet_in->NumRows()); for (int32 r = 0; r < ones.Dim(); ++r) ones(r) = 1.0f; }   // LossItf::Eval's unit frame weights
        train_step.Run(&nnet, xent, *nnet_in, ones, *nnet_tgt);
      } else {
        if (!crossvalidate) nnet.Propagate(*nnet_in, &nnet_out);
        else nnet.Feedforward(*nnet_in, &nnet_out);
        loss.Eval(nnet_out, *nnet_tgt, &obj_diff);
        if (!crossvalidate) nnet.Backpropagate(obj_diff, nullptr);
      }
      total_frames += nnet_in->NumRows();
      report_frames += nnet_in->NumRows();
      wopts.Progress(nnet_in->NumRows());
      if (report_period > 0 && report_frames >= report_period) {
        KALDI_LOG << loss.Report();
        report_frames -= report_period;
      }
    }
    wopts.Finish();
    if (!crossvalidate && wopts.WritesModel()) nnet.Write(target_model_filename, binary);
    KALDI_LOG << "[" << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << ", " << (randomize ? "RANDOMIZED" : "NOT-RANDOMIZED") << ", "
              << time.Elapsed() / 60 << " min, fps" << total_frames / time.Elapsed() << "]";
    KALDI_LOG << loss.Report();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}
