// io.h -- Kaldi-compatible token / basic-type / integer-vector stream I/O and file open helpers.
// Byte formats follow src/base/io-funcs{.h,-inl.h,.cc} and src/util/kaldi-io.cc: binary streams start
// with "\0B"; tokens are "<Tok> "; binary basic types carry a 1-byte size prefix; integer vectors are
// {char sizeof, int32 n, raw}.  Model files written here are readable by the reference and vice versa.
#ifndef ASLP_HOST_IO_H_
#define ASLP_HOST_IO_H_
#include <fstream>
#include <memory>
#include "base.h"

namespace kaldi {

void WriteToken(std::ostream& os, bool binary, const std::string& token);
void ReadToken(std::istream& is, bool binary, std::string* token);
void ExpectToken(std::istream& is, bool binary, const std::string& token);
int Peek(std::istream& is, bool binary);
int PeekToken(std::istream& is, bool binary);    // first char after '<' of the next token

void WriteBasicType(std::ostream& os, bool binary, int32 v);
void WriteBasicType(std::ostream& os, bool binary, float v);
void WriteBasicType(std::ostream& os, bool binary, double v);
void WriteBasicType(std::ostream& os, bool binary, bool v);
void ReadBasicType(std::istream& is, bool binary, int32* v);
void ReadBasicType(std::istream& is, bool binary, float* v);
void ReadBasicType(std::istream& is, bool binary, double* v);
void ReadBasicType(std::istream& is, bool binary, bool* v);

void WriteIntegerVector(std::ostream& os, bool binary, const std::vector<int32>& v);
void ReadIntegerVector(std::istream& is, bool binary, std::vector<int32>* v);

// "file", "-" (stdin/stdout) ; detects / emits the "\0B" binary header (Input/Output of util/kaldi-io.h)
class Input {
 public:
  Input() : is_(nullptr) {}
  explicit Input(const std::string& rxfilename, bool* binary = nullptr) : is_(nullptr) { Open(rxfilename, binary); }
  void Open(const std::string& rxfilename, bool* binary = nullptr);
  void OpenTextMode(const std::string& rxfilename);
  std::istream& Stream() { return *is_; }
  void Close() { file_.reset(); is_ = nullptr; }
 private:
  std::unique_ptr<std::ifstream> file_;
  std::istream* is_;
};
class Output {
 public:
  Output(const std::string& wxfilename, bool binary, bool write_header = true);
  ~Output() { Close(); }
  std::ostream& Stream() { return *os_; }
  void Close();
 private:
  std::unique_ptr<std::ofstream> file_;
  std::ostream* os_;
  std::string name_;
};

bool ConvertStringToInteger(const std::string& s, int32* out);
void SplitStringToVector(const std::string& full, const char* delim, bool omit_empty, std::vector<std::string>* out);
bool SplitStringToIntegers(const std::string& full, const char* delim, bool omit_empty, std::vector<int32>* out);

}  // namespace kaldi
#endif
