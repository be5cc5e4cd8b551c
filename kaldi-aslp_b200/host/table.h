// table.h -- the slice of Kaldi's table I/O the trainer mains use (src/util/kaldi-table.h, kaldi-holder.h):
//   SequentialBaseFloatMatrixReader   "ark:file", "scp:file" (entries "key path" or "key path:offset")
//   RandomAccessPosteriorReader       Posterior archives  (src/hmm/posterior.cc:29-99 on-disk format)
//   RandomAccessInt32VectorReader     std::vector<int32> archives (CTC label sequences)
//   BaseFloatMatrixWriter             "ark:file", "ark,t:file"
// Option letters after the type ("ark,s,cs:") are accepted and ignored.  Sources and sinks are Kaldi's extended file names
// (src/util/kaldi-io.h:60-93): a path, "-" (standard input / output), "cmd |" (read from a command, e.g. the recipes'
// "ark:copy-feats scp:... ark:- | apply-cmvn ... |") and "| cmd" (write into one); scp entries may be pipes too.
// Compressed feature matrices (CM / CM2, the default of copy-feats --compress) are decoded by Matrix::Read.
#ifndef ASLP_HOST_TABLE_H_
#define ASLP_HOST_TABLE_H_
#include <cstdio>
#include <ext/stdio_filebuf.h>
#include <fstream>
#include <map>
#include <memory>
#include "base.h"
#include "io.h"
#include "matrix.h"
#include "nnet-loss.h"

namespace kaldi {

struct TableSpec { bool scp; bool text; std::string file; };
inline TableSpec ParseSpecifier(const std::string& spec) {
  const size_t colon = spec.find(':');
  if (colon == std::string::npos) KALDI_ERR << "Invalid table specifier " << spec;
  std::vector<std::string> opts;
  SplitStringToVector(spec.substr(0, colon), ",", true, &opts);
  TableSpec t{false, false, spec.substr(colon + 1)};
  bool have_type = false;
  for (const std::string& o : opts) {
    if (o == "ark") have_type = true;
    else if (o == "scp") { t.scp = true; have_type = true; }
    else if (o == "t") t.text = true;
  }
  if (!have_type) KALDI_ERR << "Invalid table specifier " << spec << " (expected ark: or scp:)";
  while (!t.file.empty() && (t.file.back() == ' ' || t.file.back() == '\t')) t.file.pop_back();
  if (t.file.empty()) KALDI_ERR << "Invalid table specifier " << spec << " (empty file name)";
  return t;
}

// rxfilename: path | "-" | "command |"   (Input, src/util/kaldi-io.cc)
class InputStream {
 public:
  explicit InputStream(const std::string& rx) : pipe_(nullptr), is_(nullptr), name_(rx) {
    if (rx == "-") { is_ = &std::cin; return; }
    if (!rx.empty() && rx.back() == '|') {
      const std::string cmd = rx.substr(0, rx.size() - 1);
      pipe_ = popen(cmd.c_str(), "r");
      if (pipe_ == nullptr) KALDI_ERR << "Failed opening pipe for reading, command is: " << cmd;
      buf_.reset(new __gnu_cxx::stdio_filebuf<char>(pipe_, std::ios::in | std::ios::binary));
      pis_.reset(new std::istream(buf_.get()));
      is_ = pis_.get();
      return;
    }
    file_.open(rx, std::ios::in | std::ios::binary);
    if (!file_.is_open()) KALDI_ERR << "Cannot open " << rx;
    is_ = &file_;
  }
  ~InputStream() {
    pis_.reset();
    buf_.reset();
    if (pipe_ != nullptr) { const int rc = pclose(pipe_); if (rc != 0) KALDI_WARN << "Pipe " << name_ << " had nonzero return status " << rc; }
  }
  InputStream(const InputStream&) = delete;
  InputStream& operator=(const InputStream&) = delete;
  std::istream& Stream() { return *is_; }
  bool IsFile() const { return is_ == &file_; }
 private:
  std::ifstream file_;
  FILE* pipe_;
  std::unique_ptr<__gnu_cxx::stdio_filebuf<char>> buf_;
  std::unique_ptr<std::istream> pis_;
  std::istream* is_;
  std::string name_;
};
// wxfilename: path | "-" | "| command"   (Output)
class OutputStream {
 public:
  explicit OutputStream(const std::string& wx) : pipe_(nullptr), os_(nullptr), name_(wx) {
    if (wx == "-") { os_ = &std::cout; return; }
    if (!wx.empty() && wx[0] == '|') {
      const std::string cmd = wx.substr(1);
      pipe_ = popen(cmd.c_str(), "w");
      if (pipe_ == nullptr) KALDI_ERR << "Failed opening pipe for writing, command is: " << cmd;
      buf_.reset(new __gnu_cxx::stdio_filebuf<char>(pipe_, std::ios::out | std::ios::binary));
      pos_.reset(new std::ostream(buf_.get()));
      os_ = pos_.get();
      return;
    }
    file_.open(wx, std::ios::out | std::ios::binary);
    if (!file_.is_open()) KALDI_ERR << "Cannot open " << wx << " for writing";
    os_ = &file_;
  }
  ~OutputStream() {
    if (os_ != nullptr) os_->flush();
    pos_.reset();
    buf_.reset();
    if (pipe_ != nullptr) pclose(pipe_);
  }
  OutputStream(const OutputStream&) = delete;
  OutputStream& operator=(const OutputStream&) = delete;
  std::ostream& Stream() { return *os_; }
 private:
  std::ofstream file_;
  FILE* pipe_;
  std::unique_ptr<__gnu_cxx::stdio_filebuf<char>> buf_;
  std::unique_ptr<std::ostream> pos_;
  std::ostream* os_;
  std::string name_;
};

// reads the next whitespace-terminated key of an archive; false at end of file
inline bool ReadArchiveKey(std::istream& is, std::string* key) {
  key->clear();
  is >> *key;
  if (is.eof() && key->empty()) return false;
  if (is.fail()) return false;
  const int c = is.get();                       // the single separator after the key
  if (c != ' ' && c != '\t' && c != '\n') KALDI_ERR << "Invalid archive: expected space after key " << *key;
  return true;
}
inline bool ReadBinaryFlag(std::istream& is) {   // consumes "\0B" if present
  if (is.peek() == '\0') { is.get(); if (is.get() != 'B') KALDI_ERR << "Invalid binary header in archive"; return true; }
  return false;
}

struct MatrixHolder {
  typedef Matrix<BaseFloat> T;
  static void Read(std::istream& is, T* v) { const bool b = ReadBinaryFlag(is); v->Read(is, b); }
};
struct Int32VectorHolder {
  typedef std::vector<int32> T;
  static void Read(std::istream& is, T* v) {
    const bool b = ReadBinaryFlag(is);
    if (b) { ReadIntegerVector(is, true, v); return; }
    std::string line;                           // text form: integers up to end of line
    std::getline(is, line);
    if (!SplitStringToIntegers(line, " \t\r", true, v)) KALDI_ERR << "Invalid integer vector line: " << line;
  }
};
// Vector<BaseFloat> ("FV" binary or "[ ... ]" text) and plain BaseFloat values: the per-frame / per-utterance weight tables of
// the trainers (--frame-weights, --utt-weights; src/util/kaldi-holder.h KaldiObjectHolder / BasicHolder)
struct BaseFloatVectorHolder {
  typedef Vector<BaseFloat> T;
  static void Read(std::istream& is, T* v) { const bool b = ReadBinaryFlag(is); v->Read(is, b); }
};
struct BaseFloatHolder {
  typedef BaseFloat T;
  static void Read(std::istream& is, T* v) {
    const bool b = ReadBinaryFlag(is);
    ReadBasicType(is, b, v);
    if (!b) { std::string rest; std::getline(is, rest); }   // a text archive holds one value per line
  }
};
struct PosteriorHolder {
  typedef Posterior T;
  // src/hmm/posterior.cc:57-99
  static void Read(std::istream& is, T* post) {
    const bool b = ReadBinaryFlag(is);
    post->clear();
    if (b) {
      int32 sz; ReadBasicType(is, true, &sz);
      if (sz < 0 || sz > 10000000) KALDI_ERR << "Reading posterior: got negative or improbably large size " << sz;
      post->resize(sz);
      for (auto& fr : *post) {
        int32 sz2; ReadBasicType(is, true, &sz2);
        if (sz2 < 0) KALDI_ERR << "Reading posteriors: got negative size";
        fr.resize(sz2);
        for (auto& pr : fr) { ReadBasicType(is, true, &pr.first); ReadBasicType(is, true, &pr.second); }
      }
      return;
    }
    std::string line;
    std::getline(is, line);
    std::vector<std::string> tok;
    SplitStringToVector(line, " \t\r", true, &tok);
    size_t i = 0;
    while (i < tok.size()) {
      if (tok[i] != "[") KALDI_ERR << "Invalid posterior line (expected '['): " << line;
      ++i;
      std::vector<std::pair<int32, BaseFloat>> fr;
      while (i < tok.size() && tok[i] != "]") {
        if (i + 1 >= tok.size()) KALDI_ERR << "Invalid posterior line: " << line;
        int32 id;
        if (!ConvertStringToInteger(tok[i], &id)) KALDI_ERR << "Invalid posterior line: " << line;
        fr.push_back(std::make_pair(id, static_cast<BaseFloat>(std::atof(tok[i + 1].c_str()))));
        i += 2;
      }
      if (i >= tok.size()) KALDI_ERR << "Invalid posterior line (missing ']'): " << line;
      ++i;
      post->push_back(fr);
    }
  }
};

template <class Holder>
class SequentialTableReader {
 public:
  typedef typename Holder::T T;
  explicit SequentialTableReader(const std::string& rspecifier) : spec_(ParseSpecifier(rspecifier)), in_(spec_.file), main_(in_.Stream()), done_(false) {
    Next();
  }
  bool Done() const { return done_; }
  const std::string& Key() const { return key_; }
  const T& Value() const { return value_; }
  void Next() {
    if (!spec_.scp) {
      if (!ReadArchiveKey(main_, &key_)) { done_ = true; return; }
      Holder::Read(main_, &value_);
      return;
    }
    std::string line;
    while (std::getline(main_, line)) {
      // "key rxfilename": the rxfilename is the rest of the line (it contains spaces when it is a command)
      const size_t k0 = line.find_first_not_of(" \t\r");
      if (k0 == std::string::npos) continue;
      const size_t k1 = line.find_first_of(" \t", k0);
      if (k1 == std::string::npos) KALDI_ERR << "Invalid scp line: " << line;
      key_ = line.substr(k0, k1 - k0);
      size_t p0 = line.find_first_not_of(" \t", k1), p1 = line.find_last_not_of(" \t\r");
      if (p0 == std::string::npos) KALDI_ERR << "Invalid scp line: " << line;
      std::string path = line.substr(p0, p1 - p0 + 1);
      long long off = -1;
      const size_t c = path.rfind(':');
      if (c != std::string::npos && c + 1 < path.size() && path.find_first_not_of("0123456789", c + 1) == std::string::npos) {
        off = std::atoll(path.c_str() + c + 1);
        path = path.substr(0, c);
      }
      InputStream obj(path);                    // a path, or "command |" (then the line's second field runs to the end of the line)
      if (off >= 0) obj.Stream().seekg(off);
      Holder::Read(obj.Stream(), &value_);
      return;
    }
    done_ = true;
  }
 private:
  TableSpec spec_;
  InputStream in_;
  std::istream& main_;
  bool done_;
  std::string key_;
  T value_;
};

// whole table in memory: the trainers look targets up by utterance key in feature order
template <class Holder>
class RandomAccessTableReader {
 public:
  typedef typename Holder::T T;
  RandomAccessTableReader() : open_(false) {}
  explicit RandomAccessTableReader(const std::string& rspecifier) : open_(false) { Open(rspecifier); }
  bool Open(const std::string& rspecifier) {
    map_.clear();
    for (SequentialTableReader<Holder> r(rspecifier); !r.Done(); r.Next()) map_[r.Key()] = r.Value();
    open_ = true;
    return true;
  }
  bool IsOpen() const { return open_; }
  bool Close() { map_.clear(); open_ = false; return true; }
  bool HasKey(const std::string& k) const { return map_.count(k) != 0; }
  const T& Value(const std::string& k) const {
    auto it = map_.find(k);
    if (it == map_.end()) KALDI_ERR << "Value() called for key " << k << " which is not in the table";
    return it->second;
  }
 private:
  std::map<std::string, T> map_;
  bool open_;
};

typedef RandomAccessTableReader<BaseFloatVectorHolder> RandomAccessBaseFloatVectorReader;
typedef RandomAccessTableReader<BaseFloatHolder> RandomAccessBaseFloatReader;
typedef RandomAccessTableReader<MatrixHolder> RandomAccessBaseFloatMatrixReader;
typedef SequentialTableReader<MatrixHolder> SequentialBaseFloatMatrixReader;
typedef RandomAccessTableReader<PosteriorHolder> RandomAccessPosteriorReader;
typedef RandomAccessTableReader<Int32VectorHolder> RandomAccessInt32VectorReader;

class BaseFloatMatrixWriter {
 public:
  explicit BaseFloatMatrixWriter(const std::string& wspecifier) : spec_(ParseSpecifier(wspecifier)) {
    if (spec_.scp) KALDI_ERR << "scp output is not supported: " << wspecifier;
    out_.reset(new OutputStream(spec_.file));
  }
  void Write(const std::string& key, const Matrix<BaseFloat>& m) {
    std::ostream& os_ = out_->Stream();
    os_ << key << ' ';
    if (!spec_.text) { os_.put('\0'); os_.put('B'); }
    m.Write(os_, !spec_.text);
  }
 private:
  TableSpec spec_;
  std::unique_ptr<OutputStream> out_;
};

}  // namespace kaldi
#endif
