// aslp-nnet-train-lstm-streams -- multi-stream truncated-BPTT training of (projected) LSTMs with delayed targets, same
// command line, batching (SequenceDataReader), bookkeeping and log lines as
// src/aslp-nnetbin/aslp-nnet-train-lstm-streams.cc:24-240 (BASELINE config 2).
#include <memory>
#include "batch-feeder.h"
#include "nnet-nnet.h"
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
        "Perform one iteration of LSTM training by Stochastic Gradient Descent.\n"
        "This version use pdf-posterior as targets, prepared typically by ali-to-post.\n"
        "The updates are done per-utterance, shuffling options are dummy for compatibility reason.\n"
        "\n"
        "Usage:  aslp-nnet-train-lstm-streams [options] <feature-rspecifier> <targets-rspecifier> <model-in> [<model-out>]\n"
        "e.g.: \n"
        " aslp-nnet-train-lstm-streams scp:feature.scp ark:posterior.ark nnet.init nnet.iter1\n";
    ParseOptions po(usage);
    NnetTrainOptions trn_opts;
    trn_opts.Register(&po);
    NnetDataRandomizerOptions rnd_opts;
    rnd_opts.Register(&po);
    SequenceDataReaderOptions read_opts;
    read_opts.Register(&po);
    bool binary = true, crossvalidate = false;
    po.Register("binary", &binary, "Write output in binary mode");
    po.Register("cross-validate", &crossvalidate, "Perform cross-validation (don't backpropagate)");
    std::string objective_function = "xent";
    po.Register("objective-function", &objective_function, "Objective function : xent|mse");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    int32 gpu_id = -1;
    po.Register("gpu-id", &gpu_id, "selected gpu id, if negative then select automaticly");
    bool randomize = false;
    po.Register("randomize", &randomize, "Dummy option, for compatibility...");
    int32 report_period = 200;
    po.Register("report-period", &report_period, "Number of sentence for one report log, default(200)");
    int32 dump_interval = 0;
    po.Register("dump-interval", &dump_interval, "---LSTM--- num utts between model dumping [ 0 == disabled ]");
    WorkerOptions wopts;       // --worker-type: the worker of src/aslp-parallelbin/aslp-nnet-train-lstm-stream-worker.cc
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
    wopts.Create(&nnet, crossvalidate);
    long long total_frames = 0;
    int32 num_done = 0, num_sentence = 0;
    std::unique_ptr<LossItf> loss_holder;                  // LossItf* as in the reference's worker mains (xent | mse)
    if (objective_function == "xent") loss_holder.reset(new Xent);
    else if (objective_function == "mse") loss_holder.reset(new Mse);
    else KALDI_ERR << "Unsupported objective function: " << objective_function;
    LossItf& loss = *loss_holder;
    Timer time;
    KALDI_LOG << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << " STARTED";
    SequenceDataReader reader(feature_rspecifier, targets_rspecifier, read_opts);
    CuMatrix<BaseFloat> nnet_out, obj_diff, nnet_in;
    // SequenceDataReader::ReadData on the feeder thread, into a page-locked slot (same stream bookkeeping, data-reader.cc:200-324)
    struct SeqBatch {
      PinnedMatrix feat;
      bool has_feat = false;                          // false on the trailing all-masked minibatch: the previous features stay
      Posterior tgt;
      Vector<BaseFloat> frame_mask;
      std::vector<int32> new_utt_flags;
    };
    auto fill = [&](SeqBatch* b) -> bool {
      if (reader.Done()) return false;
      b->has_feat = reader.ReadDataHost(&b->feat, &b->tgt, &b->frame_mask);
      b->new_utt_flags = reader.GetNewUttFlags();
      return true;
    };
    BatchFeeder<SeqBatch> feeder(fill, /*attach_device=*/false);
    while (SeqBatch* b = feeder.Next()) {
      const std::vector<int32> new_utt_flags = b->new_utt_flags;
      nnet.ResetLstmStreams(new_utt_flags);
      if (b->has_feat) {
        nnet_in.Resize(b->feat.NumRows(), b->feat.NumCols(), kUndefined);
        nnet_in.CopyFromHost(b->feat.Data(), b->feat.Stride());       // asynchronous: the slot is page-locked
      }
      if (!crossvalidate) nnet.Propagate(nnet_in, &nnet_out);
      else nnet.Feedforward(nnet_in, &nnet_out);
      loss.Eval(b->frame_mask, nnet_out, b->tgt, &obj_diff);
      int frame_progress = 0;
      for (int32 i = 0; i < b->frame_mask.Dim(); i++) frame_progress += static_cast<int>(b->frame_mask(i));
      feeder.Release(b);                              // Xent::Eval has uploaded mask and targets
      if (!crossvalidate) { wopts.BeforeBackpropagate(frame_progress); nnet.Backpropagate(obj_diff, nullptr); }
      total_frames += frame_progress;
      wopts.Progress(frame_progress);
      int num_done_progress = 0;
      for (size_t i = 0; i < new_utt_flags.size(); i++) num_done_progress += new_utt_flags[i];
      num_done += num_done_progress;
      num_sentence += num_done_progress;
      if (num_sentence >= report_period) {
        KALDI_LOG << loss.Report();
        num_sentence -= report_period;
      }
      if (dump_interval > 0 && (num_done - num_done_progress) / dump_interval != (num_done / dump_interval) && !crossvalidate) {
        char nnet_name[512];
        snprintf(nnet_name, sizeof(nnet_name), "%s_utt%d", target_model_filename.c_str(), num_done);
        nnet.Write(nnet_name, binary);
      }
    }
    feeder.Join();
    wopts.Finish();
    if (!crossvalidate && wopts.WritesModel()) nnet.Write(target_model_filename, binary);
    KALDI_LOG << "Done " << num_done << " files, [" << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << ", "
              << (randomize ? "RANDOMIZED" : "NOT-RANDOMIZED") << ", " << time.Elapsed() / 60 << " min, fps" << total_frames / time.Elapsed() << "]";
    KALDI_LOG << loss.Report();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}
