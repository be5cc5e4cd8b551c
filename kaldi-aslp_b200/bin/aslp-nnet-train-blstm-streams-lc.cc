// aslp-nnet-train-blstm-streams-lc -- latency-controlled BLSTM multi-stream training (chunk + right context), same
// command line, stream book-keeping (curt / lent / new_utt_flags), padding rules (zero feature rows, last target
// repeated, mask 0 for the look-ahead rows) and log lines as src/aslp-nnetbin/aslp-nnet-train-blstm-streams-lc.cc:35-394.
// With --worker-type it is the worker of src/aslp-parallelbin/aslp-nnet-train-lc-blstm-streams-worker.cc:176-330
// (one process per GPU; that binary spells the look-ahead flag --right_splice: '-' and '_' are interchangeable here).
#include <memory>
#include "batch-feeder.h"
#include "nnet-nnet.h"
#include "nnet-loss.h"
#include "nnet-randomizer.h"
#include "nnet-trnopts.h"
#include "parallel-async.h"
#include "parse-options.h"
#include "worker-opts.h"
#include "table.h"

int main(int argc, char* argv[]) {
  using namespace kaldi;
  using namespace kaldi::aslp_nnet;
  try {
    const char* usage =
        "Perform one iteration of Latency Control BLSTM training by Stochastic Gradient Descent.\n"
        "This version use pdf-posterior as targets, prepared typically by ali-to-post.\n"
        "The updates are done per-utterance, shuffling options are dummy for compatibility reason.\n"
        "\n"
        "Usage: aslp-nnet-train-lstm-streams-lc [options] <feature-rspecifier> <targets-rspecifier> <model-in> [<model-out>]\n"
        "e.g.: \n"
        " aslp-nnet-train-lstm-streams-lc scp:feature.scp ark:posterior.ark nnet.init nnet.iter1\n";
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
    std::string use_gpu = "yes";
    po.Register("use-gpu", &use_gpu, "yes|no|optional, only has effect if compiled with CUDA");
    int32 chunk_size = 64;
    po.Register("chunk-size", &chunk_size, "---BLSTM--- Latency-controlled BPTT chunk size");
    int32 right_splice = 16;
    po.Register("right-splice", &right_splice, "---BLSTM--- Latency-controlled BPTT right context size");
    int32 num_stream = 4;
    po.Register("num-stream", &num_stream, "---LSTM--- BPTT multi-stream training");
    int32 dump_interval = 0;
    po.Register("dump-interval", &dump_interval, "---LSTM--- num utts between model dumping [ 0 == disabled ]");
    NnetDataRandomizerOptions rnd_opts;      // dummy, for compatibility with the standard scripts
    rnd_opts.Register(&po);
    bool randomize = false;
    po.Register("randomize", &randomize, "Dummy option, for compatibility...");
    int32 report_period = 200;
    po.Register("report-period", &report_period, "Number of sentence for one report log, default(200)");
    int32 drop_len = 0;
    po.Register("drop-len", &drop_len, "if Sentence frame length greater than drop_len,then drop it, default(0, no drop)");
    // worker extension
    std::string worker_type = "";
    po.Register("worker-type", &worker_type, "Worker type(bsp | bmuf | sod | easgd | asgd); empty: single process");
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
    OptimizerOption optimizer_opts;
    optimizer_opts.Register(&po);
    po.Read(argc, argv);
    const int32 batch_size = chunk_size + right_splice;
    if (po.NumArgs() != 4 - (crossvalidate ? 1 : 0)) { po.PrintUsage(); return 1; }
    const std::string feature_rspecifier = po.GetArg(1), targets_rspecifier = po.GetArg(2), model_filename = po.GetArg(3);
    std::string target_model_filename;
    if (!crossvalidate) target_model_filename = po.GetArg(4);
    if (use_gpu == "no") KALDI_ERR << "--use-gpu=no: this build has no CPU path";
    if (!worker_type.empty()) {
      const char* lr = std::getenv("LOCAL_RANK");
      if (lr != nullptr) ASLP_OK(aslp_set_device(std::atoi(lr)));
    }

    Nnet nnet_transf;
    if (feature_transform != "") nnet_transf.Read(feature_transform);
    Nnet nnet;
    nnet.Read(model_filename);
    nnet.SetTrainOptions(trn_opts);
    nnet.SetChunkSize(chunk_size);

    std::unique_ptr<IWorker> worker;
    SyncCounter sync;                                      // worker-opts.h: frame counter + Synchronize, blocking or pipelined by layer
    if (!worker_type.empty() && !crossvalidate) {
      WorkerBootstrap boot;
      if (worker_type == "bsp") worker.reset(new BspWorker(boot.id, boot.nranks, boot.rank));
      else if (worker_type == "bmuf") worker.reset(new BmufWorker(boot.id, boot.nranks, boot.rank, bmuf_momentum, bmuf_learn_rate));
      else if (worker_type == "sod") worker.reset(new SodWorker(boot.id, boot.nranks, boot.rank, optimizer_opts));
      else if (worker_type == "easgd") worker.reset(new EasgdWorker(boot.id, boot.nranks, boot.rank, alpha));     // rank 0 runs aslp-nnet-train-server
      else if (worker_type == "asgd") worker.reset(new AsgdWorker(boot.id, boot.nranks, boot.rank));
      else KALDI_ERR << "Unsupported worker type: " << worker_type;
      sync.Attach(worker.get(), &nnet, sync_period, pipeline_sync);
    }

    long long total_frames = 0;
    SequentialBaseFloatMatrixReader feature_reader(feature_rspecifier);
    RandomAccessPosteriorReader target_reader(targets_rspecifier);
    std::unique_ptr<LossItf> loss_holder;                  // LossItf* as in the reference's worker mains (xent | mse)
    if (objective_function == "xent") loss_holder.reset(new Xent);
    else if (objective_function == "mse") loss_holder.reset(new Mse);
    else KALDI_ERR << "Unsupported objective function: " << objective_function;
    LossItf& xent = *loss_holder;
    Timer time;
    KALDI_LOG << (crossvalidate ? "CROSS-VALIDATION" : "TRAINING") << " STARTED";
    int32 num_done = 0, num_no_tgt_mat = 0, num_other_error = 0, num_sentence = 0;

    std::vector<std::string> keys(num_stream);
    std::vector<Matrix<BaseFloat>> feats(num_stream);
    std::vector<Posterior> targets(num_stream);
    std::vector<int32> curt(num_stream, 0), lent(num_stream, 0);
    const int32 feat_dim = nnet.InputDim();
    CuMatrix<BaseFloat> feat_dev, nnet_out, obj_diff;

    // One chunk minibatch, built by the feeder thread with the reference's own stream bookkeeping (:185-268): streams whose
    // utterance is used up take the next readable one, then batch_size rows per stream are packed into a page-locked slot.
    struct LcBatch {
      PinnedMatrix feat;
      Vector<BaseFloat> frame_mask;
      Posterior target;
      std::vector<int32> new_utt_flags;
    };
    CuMatrix<BaseFloat> transf_in, transf_out;                 // the feeder thread's own device buffers (its stream)
    auto fill = [&](LcBatch* b) -> bool {
      b->new_utt_flags.assign(num_stream, 0);
      for (int32 s = 0; s < num_stream; s++) {
        if (curt[s] < lent[s]) { b->new_utt_flags[s] = 0; continue; }
        while (!feature_reader.Done()) {
          const std::string key = feature_reader.Key();
          const Matrix<BaseFloat>& mat = feature_reader.Value();
          if (drop_len > 0 && mat.NumRows() > drop_len) {
            KALDI_WARN << key << ", too long, droped";
            feature_reader.Next();
            continue;
          }
          Matrix<BaseFloat> transformed;
          if (nnet_transf.NumComponents() > 0) {
            transf_in = mat;
            nnet_transf.Feedforward(transf_in, &transf_out);
            transf_out.CopyToMat(&transformed);
          } else {
            transformed = mat;
          }
          if (!target_reader.HasKey(key)) {
            KALDI_WARN << key << ", missing targets";
            num_no_tgt_mat++;                      // these two counters are the feeder thread's until feeder.Join()
            feature_reader.Next();
            continue;
          }
          const Posterior& tgt = target_reader.Value(key);
          if (transformed.NumRows() != static_cast<int32>(tgt.size())) {
            KALDI_WARN << key << ", length miss-match between feats and targets, skip";
            num_other_error++;
            feature_reader.Next();
            continue;
          }
          keys[s] = key;
          feats[s] = transformed;
          targets[s] = tgt;
          curt[s] = 0;
          lent[s] = feats[s].NumRows();
          b->new_utt_flags[s] = 1;
          feature_reader.Next();
          break;
        }
      }
      int done = 1;
      for (int32 s = 0; s < num_stream; s++) if (curt[s] < lent[s]) done = 0;
      if (done) return false;

      b->frame_mask.Resize(batch_size * num_stream);
      b->target.resize(batch_size * num_stream);
      b->feat.Resize(batch_size * num_stream, feat_dim, kUndefined);
      for (int32 t = 0; t < batch_size; t++) {
        for (int32 s = 0; s < num_stream; s++) {
          const int32 row = t * num_stream + s;
          if (curt[s] < lent[s]) {
            b->frame_mask(row) = (t >= chunk_size) ? 0.0f : 1.0f;
            b->target[row] = targets[s][curt[s]];
            std::copy(feats[s].RowData(curt[s]), feats[s].RowData(curt[s]) + feat_dim, b->feat.RowData(row));
          } else {
            b->frame_mask(row) = 0.0f;
            if (lent[s] > 0) b->target[row] = targets[s][lent[s] - 1]; else b->target[row].clear();
            std::fill(b->feat.RowData(row), b->feat.RowData(row) + feat_dim, 0.0f);      // zero frames, not the last frame (:261)
          }
          curt[s]++;
        }
      }
      for (int32 s = 0; s < num_stream; s++) curt[s] = curt[s] - right_splice;
      return true;
    };
    BatchFeeder<LcBatch> feeder(fill, /*attach_device=*/nnet_transf.NumComponents() > 0);

    while (LcBatch* b = feeder.Next()) {
      nnet.ResetLstmStreams(b->new_utt_flags);
      feat_dev.Resize(b->feat.NumRows(), feat_dim, kUndefined);
      feat_dev.CopyFromHost(b->feat.Data(), b->feat.Stride());           // asynchronous: the slot is page-locked
      if (!crossvalidate) nnet.Propagate(feat_dev, &nnet_out);
      else nnet.Feedforward(feat_dev, &nnet_out);
      xent.Eval(b->frame_mask, nnet_out, b->target, &obj_diff);

      int frame_progress = 0;
      for (int32 i = 0; i < b->frame_mask.Dim(); i++) frame_progress += static_cast<int>(b->frame_mask(i));
      int num_done_progress = 0;
      for (size_t i = 0; i < b->new_utt_flags.size(); i++) num_done_progress += b->new_utt_flags[i];
      feeder.Release(b);                              // Xent::Eval has uploaded mask and targets (pageable: staged at the call)
      if (!crossvalidate) { sync.BeforeBackpropagate(frame_progress); nnet.Backpropagate(obj_diff, nullptr); }

      total_frames += frame_progress;
      num_done += num_done_progress;
      num_sentence += num_done_progress;
      if (num_sentence >= report_period) {
        KALDI_LOG << xent.Report();
        num_sentence -= report_period;
      }
      sync.Progress(frame_progress);
      if (dump_interval > 0 && (num_done - num_done_progress) / dump_interval != (num_done / dump_interval) && !crossvalidate) {
        char nnet_name[512];
        snprintf(nnet_name, sizeof(nnet_name), "%s_utt%d", target_model_filename.c_str(), num_done);
        nnet.Write(nnet_name, binary);
      }
    }
    feeder.Join();
    if (worker) FinishWorker(worker.get(), &nnet);        // Stop(), then the BatchNorm statistics of all ranks (worker-opts.h)
    if (!crossvalidate && (!worker || worker->IsMainNode())) nnet.Write(target_model_filename, binary);
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
