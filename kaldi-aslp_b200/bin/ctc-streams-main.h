// ctc-streams-main.h -- the body shared by aslp-nnet-train-warp-ctc-streams and aslp-nnet-train-ctc-streams: the two reference
// mains (src/aslp-nnetbin/aslp-nnet-train-warp-ctc-streams.cc:21-235, aslp-nnet-train-ctc-streams.cc) differ only in the loss
// class (WarpCtc::Eval / ErrorRate vs Ctc::EvalParallel / ErrorRateMSeq).  Multi-stream whole-utterance CTC training: same
// command line, batching rule (num_stream utterances or frame_limit padded frames), learn-rate normalisation and log lines.
// Added for BASELINE config 5 (the reference has no CTC worker binary): --worker-type=bsp|bmuf with the worker flags of
// src/aslp-parallelbin/aslp-nnet-train-lc-blstm-streams-worker.cc:60-100 turns the same loop into one rank of a
// data-parallel job (one process per GPU; RANK / WORLD_SIZE / LOCAL_RANK from the environment).
#ifndef ASLP_BIN_CTC_STREAMS_MAIN_H_
#define ASLP_BIN_CTC_STREAMS_MAIN_H_
#include "batch-feeder.h"
#include "nnet-nnet.h"
#include "nnet-loss.h"
#include "nnet-randomizer.h"
#include "nnet-trnopts.h"
#include "parallel-async.h"
#include "parse-options.h"
#include "worker-opts.h"
#include "table.h"

namespace kaldi {
namespace aslp_nnet {

// LossAdapter: { void SetReportStep(int); void Eval(keys, frames, net_out, labels, &diff); void ErrorRate(frames, net_out, labels); std::string Report(); }
template <class LossAdapter>
int CtcStreamsMain(int argc, char* argv[], const char* usage) {
  try {
    ParseOptions po(usage);
    NnetTrainOptions trn_opts;
    trn_opts.Register(&po);
    bool binary = true, crossvalidate = false;
    po.Register("binary", &binary, "Write model  in binary mode");
    po.Register("cross-validate", &crossvalidate, "Perform cross-validation (no backpropagation)");
    int32 num_stream = 5;
    po.Register("num-stream", &num_stream, "Number of sequences processed in parallel");
    double frame_limit = 100000;
    po.Register("frame-limit", &frame_limit, "Max number of frames to be processed");
    NnetDataRandomizerOptions rnd_opts;      // dummy, for compatibility with the standard scripts
    rnd_opts.Register(&po);
    bool randomize = false;
    po.Register("randomize", &randomize, "Dummy option, for compatibility...");
    int32 report_step = 100;
    po.Register("report-step", &report_step, "Step (number of sequences) for status reporting");
    int32 report_period = 200;
    po.Register("report-period", &report_period, "Number of sentence for one report log, default(200)");
    int32 drop_len = 0;
    po.Register("drop-len", &drop_len, "if Sentence frame length greater than drop_len,then drop it, default(0, no drop)");
    int32 skip_width = 0;
    po.Register("skip-width", &skip_width, "num of frame for one skip(default 0, not use skip)");
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    // data-parallel extension (config 5)
    std::string worker_type = "";
    po.Register("worker-type", &worker_type, "Data-parallel worker: bsp | bmuf | easgd | asgd (empty: single process)");
    float alpha = 0.5f;
    po.Register("alpha", &alpha, "Moving rate alpha for easgd worker");
    int32 sync_period = 25600;
    po.Register("sync-period", &sync_period, "number of frames for one sync with other workers");
    bool pipeline_sync = true;
    po.Register("pipeline-sync", &pipeline_sync, "bsp | bmuf | sod: exchange each layer's tensors behind its Update, under the backward pass "
                "(same result as the blocking exchange after the minibatch; every rank must use the same setting)");
    float bmuf_momentum = 0.9f, bmuf_learn_rate = 1.0f;
    po.Register("bmuf-momentum", &bmuf_momentum, "bmuf block momentum");
    po.Register("bmuf-learn-rate", &bmuf_learn_rate, "bmuf block learning rate");
    po.Read(argc, argv);
    if (po.NumArgs() != 4 - (crossvalidate ? 1 : 0)) { po.PrintUsage(); return 1; }
    const std::string feature_rspecifier = po.GetArg(1), targets_rspecifier = po.GetArg(2), model_filename = po.GetArg(3);
    std::string target_model_filename;
    if (!crossvalidate) target_model_filename = po.GetArg(4);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (!worker_type.empty()) {
      const char* lr = std::getenv("LOCAL_RANK");
      if (lr != nullptr) ASLP_OK(aslp_set_device(std::atoi(lr)));
    }

    Nnet net;
    net.Read(model_filename);
    net.SetTrainOptions(trn_opts);
    const float norm_lr = trn_opts.learn_rate;

    std::unique_ptr<IWorker> worker;
    SyncCounter sync;                                      // worker-opts.h: frame counter + Synchronize, blocking or pipelined by layer
    if (!worker_type.empty() && !crossvalidate) {
      WorkerBootstrap boot;
      if (worker_type == "bsp") worker.reset(new BspWorker(boot.id, boot.nranks, boot.rank));
      else if (worker_type == "bmuf") worker.reset(new BmufWorker(boot.id, boot.nranks, boot.rank, bmuf_momentum, bmuf_learn_rate));
      else if (worker_type == "easgd") worker.reset(new EasgdWorker(boot.id, boot.nranks, boot.rank, alpha));     // rank 0 runs aslp-nnet-train-server
      else if (worker_type == "asgd") worker.reset(new AsgdWorker(boot.id, boot.nranks, boot.rank));
      else KALDI_ERR << "Unsupported worker type: " << worker_type;
      sync.Attach(worker.get(), &net, sync_period, pipeline_sync);
    }

    long long total_frames = 0;
    SequentialBaseFloatMatrixReader feature_reader(feature_rspecifier);
    RandomAccessInt32VectorReader targets_reader(targets_rspecifier);
    LossAdapter ctc;
    ctc.SetReportStep(report_step);
    CuMatrix<BaseFloat> net_out, obj_diff;
    Timer time;
    KALDI_LOG << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << " STARTED";
    const int32 feat_dim = net.InputDim();
    int32 num_done = 0, num_no_tgt_mat = 0, num_other_error = 0, num_sentence = 0;
    CuMatrix<BaseFloat> feat_mat_dev;

    // One group of utterances, read / filtered / packed by the feeder thread exactly as the reference's loop does it
    // (aslp-nnet-train-warp-ctc-streams.cc:116-175), into a page-locked slot.
    struct CtcBatch {
      PinnedMatrix feat;                              // stream-interleaved, zero padded to the longest utterance of the group
      std::vector<int32> frame_num_utt;
      std::vector<std::string> keys;
      std::vector<std::vector<int32>> labels;
      int32 num_valid_frame = 0;
      bool last = false;                              // the feature reader was exhausted when the group closed
    };
    std::vector<Matrix<BaseFloat>> feats_utt(num_stream);
    auto fill = [&](CtcBatch* b) -> bool {
      b->frame_num_utt.clear(); b->keys.clear(); b->labels.clear();
      b->num_valid_frame = 0;
      int32 sequence_index = 0, max_frame_num = 0;
      for (; !feature_reader.Done(); feature_reader.Next()) {
        const std::string utt = feature_reader.Key();
        if (!targets_reader.HasKey(utt)) {
          KALDI_WARN << utt << ", missing targets";
          num_no_tgt_mat++;                          // owned by the feeder thread until feeder.Join()
          continue;
        }
        const Matrix<BaseFloat>& raw_mat = feature_reader.Value();
        if (drop_len > 0 && raw_mat.NumRows() > drop_len) {
          KALDI_WARN << utt << ", too long, droped";
          continue;
        }
        Matrix<BaseFloat>& mat = feats_utt[sequence_index];
        if (skip_width > 1) {
          const int32 skip_len = (raw_mat.NumRows() - 1) / skip_width + 1;
          mat.Resize(skip_len, raw_mat.NumCols());
          for (int32 i = 0; i < skip_len; i++)
            std::copy(raw_mat.RowData(i * skip_width), raw_mat.RowData(i * skip_width) + raw_mat.NumCols(), mat.RowData(i));
        } else {
          mat = raw_mat;
        }
        if (max_frame_num < mat.NumRows()) max_frame_num = mat.NumRows();
        b->labels.push_back(targets_reader.Value(utt));
        b->keys.push_back(utt);
        b->frame_num_utt.push_back(mat.NumRows());
        sequence_index++;
        if (static_cast<int32>(b->frame_num_utt.size()) == num_stream || b->frame_num_utt.size() * max_frame_num > frame_limit) {
          feature_reader.Next();
          break;
        }
      }
      const int32 cur_sequence_num = static_cast<int32>(b->frame_num_utt.size());
      b->last = feature_reader.Done();
      if (cur_sequence_num == 0) return false;  // nothing left (the reference would assert inside Propagate here)
      b->feat.Resize(cur_sequence_num * max_frame_num, feat_dim, kSetZero);
      for (int32 s = 0; s < cur_sequence_num; s++) {
        const Matrix<BaseFloat>& m = feats_utt[s];
        KALDI_ASSERT(m.NumCols() == feat_dim);
        for (int32 r = 0; r < b->frame_num_utt[s]; r++)
          std::copy(m.RowData(r), m.RowData(r) + feat_dim, b->feat.RowData(r * cur_sequence_num + s));
        b->num_valid_frame += b->frame_num_utt[s];
      }
      return true;
    };
    BatchFeeder<CtcBatch> feeder(fill, /*attach_device=*/false);

    while (CtcBatch* b = feeder.Next()) {
      const int32 cur_sequence_num = static_cast<int32>(b->frame_num_utt.size());
      const int32 num_valid_frame = b->num_valid_frame, batch_rows = b->feat.NumRows();
      net.SetSeqLengths(b->frame_num_utt);
      trn_opts.learn_rate = norm_lr / num_valid_frame;        // per-minibatch learn-rate normalisation (:177)
      net.SetTrainOptions(trn_opts);
      feat_mat_dev.Resize(b->feat.NumRows(), feat_dim, kUndefined);
      feat_mat_dev.CopyFromHost(b->feat.Data(), b->feat.Stride());      // asynchronous: the slot is page-locked
      // the loss below needs this group's keys / lengths / labels after the slot has gone back to the feeder
      const std::vector<int32> frame_num_utt = b->frame_num_utt;
      const std::vector<std::string> keys = b->keys;
      std::vector<std::vector<int32>> labels = b->labels;     // (the Eesen front end takes it non-const, as the reference does)
      const bool last = b->last;
      feeder.Release(b);
      if (!crossvalidate) net.Propagate(feat_mat_dev, &net_out);
      else net.Feedforward(feat_mat_dev, &net_out);
      ctc.Eval(keys, frame_num_utt, net_out, labels, &obj_diff);
      ctc.ErrorRate(frame_num_utt, net_out, labels);
      if (!crossvalidate) { sync.BeforeBackpropagate(num_valid_frame); net.Backpropagate(obj_diff, nullptr); }
      num_done += cur_sequence_num;
      total_frames += batch_rows;
      num_sentence += cur_sequence_num;
      if (num_sentence >= report_period) {
        KALDI_LOG << ctc.Report();
        num_sentence -= report_period;
      }
      sync.Progress(num_valid_frame);
      if (last) break;
    }
    feeder.Join();
    if (worker) FinishWorker(worker.get(), &net);        // Stop(), then the BatchNorm statistics of all ranks (worker-opts.h)
    if (!crossvalidate) KALDI_LOG << net.InfoGradient();
    if (!crossvalidate && (!worker || worker->IsMainNode())) net.Write(target_model_filename, binary);
    KALDI_LOG << "Done " << num_done << " files, " << num_no_tgt_mat << " with no targets, " << num_other_error << " with other errors. "
              << "[" << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << ", " << time.Elapsed() / 60 << " min, fps"
              << total_frames / time.Elapsed() << "]";
    KALDI_LOG << ctc.Report();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}

}  // namespace aslp_nnet
}  // namespace kaldi
#endif
