// ref_driver.cc -- TEST / BASELINE INFRASTRUCTURE.  Drives the UNMODIFIED reference CPU classes
// (compiled from /root/reference/src by oracle/Makefile with HAVE_CUDA=0) on given inputs and dumps what the
// parity tests compare: net output, input derivative, per-component buffers, parameters after the update, loss.
// It is our own code (no reference source is copied): it only *calls* kaldi::aslp_nnet::Nnet, Xent, WarpCtc.
//
//   ref_driver init  <proto> <model-out> <seed> <binary 0|1>
//   ref_driver step  <model-in> <spec-file> <out-dir>
//   ref_driver bench <model-in> <spec-file>            (times `iters` training minibatches, prints one JSON line)
//   ref_driver compress <matrix-in> <archive-out> <key> <1|2>   (the reference's CompressedMatrix writer, format CM or CM2: fixtures for the I/O tests)
//
// spec-file: one `key value` per line
//   input <kaldi matrix file>          features [rows, dim], stream-interleaved for recurrent nets
//   out_diff <kaldi matrix file>       explicit d(loss)/d(output)            (loss none)
//   loss none|xent|mse|ctc                mse: Mse::Eval on the same posterior targets as xent
//   srand n                            std::srand(n) before the first iteration (Dropout draws its masks from rand())
//   targets <file>                     xent: one int per row (text, whitespace separated)
//   frame_mask <file>                  xent: one float per row (optional)
//   labels <file>                      ctc: one utterance per line, ints
//   seq_lengths a,b,c                  SetSeqLengths (also forwarded to BLstmProjectedStreamsLC components directly:
//                                      the reference's Nnet::SetSeqLengths omits that type, nnet-nnet.cc:492-523)
//   reset_flags 1,1,0                  ResetLstmStreams before every iteration
//   chunk_size n                       SetChunkSize
//   learn_rate / momentum / l2 / l1    NnetTrainOptions
//   norm_learn_rate x                  ctc: learn_rate = x / valid frames per minibatch
//   iters n                            repeat the same minibatch n times (exercises momentum + carried state)
//   dump_components 0|1
#include <sys/stat.h>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "base/kaldi-common.h"
#include "util/common-utils.h"
#include "aslp-cudamatrix/cu-matrix.h"
#include "aslp-cudamatrix/cu-vector.h"
#include "aslp-nnet/nnet-component.h"
#include "aslp-nnet/nnet-trnopts.h"
// the graph executor's live buffers (output_buf_, output_diff_buf_) are private; the public accessors return the legacy
// never-filled ones.  Every std / kaldi header is already included above, so this only opens class Nnet.
#define private public
#include "aslp-nnet/nnet-nnet.h"
#undef private
#include "aslp-nnet/nnet-blstm-projected-streams-lc.h"
#include "aslp-nnet/nnet-loss.h"
#include "aslp-nnet/warp-ctc.h"
#include "base/timer.h"
#include "matrix/compressed-matrix.h"

using namespace kaldi;
using namespace kaldi::aslp_nnet;

static std::map<std::string, std::string> ReadSpec(const std::string& f) {
  std::map<std::string, std::string> m;
  std::ifstream is(f.c_str());
  std::string line;
  while (std::getline(is, line)) {
    std::istringstream ls(line);
    std::string k, v;
    ls >> k;
    std::getline(ls, v);
    size_t a = v.find_first_not_of(" \t");
    if (k.empty() || a == std::string::npos) continue;
    m[k] = v.substr(a);
  }
  return m;
}
static std::vector<int32> Ints(const std::string& s) {
  std::vector<int32> v;
  std::string t(s);
  for (char& c : t) if (c == ',') c = ' ';
  std::istringstream is(t);
  int x;
  while (is >> x) v.push_back(x);
  return v;
}
static void ReadMat(const std::string& f, Matrix<BaseFloat>* m) { bool b; Input in(f, &b); m->Read(in.Stream(), b); }
static void WriteMat(const std::string& f, const MatrixBase<BaseFloat>& m) { Output o(f, true); m.Write(o.Stream(), true); }
static void WriteVec(const std::string& f, const VectorBase<BaseFloat>& v) { Output o(f, true); v.Write(o.Stream(), true); }

int main(int argc, char** argv) {
  try {
    if (argc < 2) { std::cerr << "usage: ref_driver init|step|bench ...\n"; return 1; }
    const std::string cmd = argv[1];
    if (cmd == "init") {
      if (argc != 6) { std::cerr << "ref_driver init <proto> <model-out> <seed> <binary>\n"; return 1; }
      std::srand(atoi(argv[4]));
      Nnet nnet;
      nnet.Init(argv[2]);
      nnet.Write(argv[3], atoi(argv[5]) != 0);
      return 0;
    }
    if (cmd == "compress") {
      if (argc != 6) { std::cerr << "ref_driver compress <matrix-in> <archive-out> <key> <1|2>\n"; return 1; }
      Matrix<BaseFloat> m;
      ReadMat(argv[2], &m);
      CompressedMatrix cm;
      if (atoi(argv[5]) == 2) {                 // format 2 has no public selector in this version: it is what a matrix with fewer than 8 rows gets
        KALDI_ASSERT(m.NumRows() < 8);
      }
      cm.CopyFromMat(m);
      Output ko(argv[3], true, false);          // binary, no header: an archive entry "key \0B<matrix>"
      ko.Stream() << argv[4] << ' ';
      ko.Stream().put('\0'); ko.Stream().put('B');
      cm.Write(ko.Stream(), true);
      Matrix<BaseFloat> back(cm.NumRows(), cm.NumCols());
      cm.CopyToMat(&back);
      WriteMat(std::string(argv[3]) + ".decoded", back);
      return 0;
    }
    if (cmd != "step" && cmd != "bench") { std::cerr << "unknown command " << cmd << "\n"; return 1; }
    const bool bench = cmd == "bench";
    std::map<std::string, std::string> sp = ReadSpec(argv[3]);
    const std::string outdir = bench ? "" : argv[4];
    if (!bench) mkdir(outdir.c_str(), 0755);
    Nnet nnet;
    nnet.Read(argv[2]);
    NnetTrainOptions opts;
    opts.learn_rate = sp.count("learn_rate") ? atof(sp["learn_rate"].c_str()) : 0.0;
    opts.momentum = sp.count("momentum") ? atof(sp["momentum"].c_str()) : 0.0;
    opts.l2_penalty = sp.count("l2") ? atof(sp["l2"].c_str()) : 0.0;
    opts.l1_penalty = sp.count("l1") ? atof(sp["l1"].c_str()) : 0.0;
    nnet.SetTrainOptions(opts);
    if (sp.count("chunk_size")) nnet.SetChunkSize(atoi(sp["chunk_size"].c_str()));
    Matrix<BaseFloat> in_h;
    ReadMat(sp["input"], &in_h);
    const std::string loss = sp.count("loss") ? sp["loss"] : "none";
    const int iters = sp.count("iters") ? atoi(sp["iters"].c_str()) : 1;
    const int warmup = sp.count("warmup") ? atoi(sp["warmup"].c_str()) : 0;      // bench: untimed leading iterations
    const bool dump_comp = sp.count("dump_components") && atoi(sp["dump_components"].c_str()) != 0;
    const float norm_lr = sp.count("norm_learn_rate") ? atof(sp["norm_learn_rate"].c_str()) : 0.0f;
    std::vector<int32> seq_lengths = sp.count("seq_lengths") ? Ints(sp["seq_lengths"]) : std::vector<int32>();
    std::vector<int32> reset_flags = sp.count("reset_flags") ? Ints(sp["reset_flags"]) : std::vector<int32>();

    Matrix<BaseFloat> od_h;
    if (loss == "none" && sp.count("out_diff")) ReadMat(sp["out_diff"], &od_h);
    Posterior post;
    Vector<BaseFloat> frame_mask;
    std::vector<std::vector<int32> > labels;
    if (loss == "xent" || loss == "mse") {
      std::ifstream ts(sp["targets"].c_str());
      int t;
      while (ts >> t) { post.push_back(std::vector<std::pair<int32, BaseFloat> >(1, std::make_pair(t, 1.0f))); }
      frame_mask.Resize(post.size());
      frame_mask.Set(1.0);
      if (sp.count("frame_mask")) { std::ifstream ms(sp["frame_mask"].c_str()); for (int i = 0; i < frame_mask.Dim(); ++i) ms >> frame_mask(i); }
    } else if (loss == "ctc") {
      std::ifstream ls(sp["labels"].c_str());
      std::string line;
      while (std::getline(ls, line)) { if (line.find_first_not_of(" \t\r") != std::string::npos) labels.push_back(Ints(line)); }
    }
    Xent xent;
    Mse mse;
    WarpCtc ctc;
    ctc.SetUseGpu(false);
    std::vector<std::string> keys;
    for (size_t i = 0; i < seq_lengths.size(); ++i) { std::ostringstream k; k << "utt" << i; keys.push_back(k.str()); }

    if (sp.count("srand")) std::srand(atoi(sp["srand"].c_str()));
    CuMatrix<BaseFloat> in(in_h), out, diff, in_diff;
    Timer timer;
    double frames_done = 0;
    for (int it = 0; it < iters; ++it) {
      if (bench && it == warmup) { timer.Reset(); frames_done = 0; }
      if (!seq_lengths.empty()) {
        nnet.SetSeqLengths(seq_lengths);
        for (int c = 0; c < nnet.NumComponents(); ++c)
          if (nnet.GetComponent(c).GetType() == Component::kBLstmProjectedStreamsLC)
            dynamic_cast<BLstmProjectedStreamsLC&>(nnet.GetComponent(c)).SetSeqLengths(seq_lengths);
      }
      if (!reset_flags.empty()) nnet.ResetLstmStreams(reset_flags);
      if (loss == "ctc" && norm_lr > 0) {
        int valid = 0;
        for (size_t i = 0; i < seq_lengths.size(); ++i) valid += seq_lengths[i];
        NnetTrainOptions o = nnet.GetTrainOptions();
        o.learn_rate = norm_lr / valid;
        nnet.SetTrainOptions(o);
      }
      nnet.Propagate(in, &out);
      if (loss == "xent") xent.Eval(frame_mask, out, post, &diff);
      else if (loss == "mse") mse.Eval(frame_mask, out, post, &diff);
      else if (loss == "ctc") ctc.Eval(keys, seq_lengths, out, labels, &diff);
      else diff = CuMatrix<BaseFloat>(od_h);
      nnet.Backpropagate(diff, &in_diff);
      frames_done += in.NumRows();
      if (bench) continue;
      std::ostringstream tag; tag << ".iter" << it;
      Matrix<BaseFloat> tmp(out.NumRows(), out.NumCols());
      out.CopyToMat(&tmp); WriteMat(outdir + "/out" + tag.str(), tmp);
      tmp.Resize(diff.NumRows(), diff.NumCols()); diff.CopyToMat(&tmp); WriteMat(outdir + "/loss_diff" + tag.str(), tmp);
      tmp.Resize(in_diff.NumRows(), in_diff.NumCols()); in_diff.CopyToMat(&tmp); WriteMat(outdir + "/in_diff" + tag.str(), tmp);
      Vector<BaseFloat> params;
      nnet.GetParams(&params);
      WriteVec(outdir + "/params" + tag.str(), params);
      if (dump_comp) {
        for (int c = 0; c < nnet.NumComponents(); ++c) {
          std::ostringstream n1, n2;
          n1 << outdir << "/comp" << c << "_out" << tag.str();
          n2 << outdir << "/comp" << c << "_out_diff" << tag.str();
          Matrix<BaseFloat> a(nnet.output_buf_[c].NumRows(), nnet.output_buf_[c].NumCols());
          nnet.output_buf_[c].CopyToMat(&a); WriteMat(n1.str(), a);
          Matrix<BaseFloat> b(nnet.output_diff_buf_[c].NumRows(), nnet.output_diff_buf_[c].NumCols());
          nnet.output_diff_buf_[c].CopyToMat(&b); WriteMat(n2.str(), b);
        }
      }
    }
    const double el = timer.Elapsed();
    if (bench) {
      std::cout << "{\"impl\": \"reference-cpu\", \"iters\": " << (iters - warmup) << ", \"frames\": " << frames_done << ", \"seconds\": " << el
                << ", \"frames_per_sec\": " << frames_done / el << "}" << std::endl;
    } else {
      std::ofstream rep((outdir + "/report.txt").c_str());
      if (loss == "xent") rep << xent.Report();
      if (loss == "mse") rep << mse.Report();
      if (loss == "ctc") rep << ctc.Report() << "\n";
    }
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return -1;
  }
}
