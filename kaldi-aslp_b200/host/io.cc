#include "io.h"
#include <cctype>
#include <cstring>

namespace kaldi {

int g_kaldi_verbose_level = 0;

void WriteToken(std::ostream& os, bool, const std::string& token) {
  os << token << " ";
  if (os.fail()) KALDI_ERR << "Write failure in WriteToken.";
}

int Peek(std::istream& is, bool binary) {
  if (!binary) is >> std::ws;
  return is.peek();
}

void ReadToken(std::istream& is, bool binary, std::string* str) {
  if (!binary) is >> std::ws;
  is >> *str;
  if (is.fail()) KALDI_ERR << "ReadToken, failed to read token at file position " << is.tellg();
  if (!isspace(is.peek())) KALDI_ERR << "ReadToken, expected space after token, saw instead " << static_cast<char>(is.peek());
  is.get();   // consume the space
}

int PeekToken(std::istream& is, bool binary) {
  if (!binary) is >> std::ws;
  bool read_bracket = false;
  if (static_cast<char>(is.peek()) == '<') { read_bracket = true; is.get(); }
  const int ans = is.peek();
  if (read_bracket) is.unget();
  return ans;
}

void ExpectToken(std::istream& is, bool binary, const std::string& token) {
  std::string got;
  const auto pos = is.tellg();
  if (!binary) is >> std::ws;
  is >> got;
  is.get();
  if (is.fail()) KALDI_ERR << "Failed to read token [started at file position " << pos << "], expected " << token;
  if (got != token) KALDI_ERR << "Expected token \"" << token << "\", got instead \"" << got << "\".";
}

template <class T> static void WriteBin(std::ostream& os, T v) {
  const char sz = static_cast<char>(sizeof(T));
  os.put(sz);
  os.write(reinterpret_cast<const char*>(&v), sizeof(T));
}
template <class T> static void ReadBin(std::istream& is, T* v, bool is_float) {
  const int sz = is.get();
  if (sz != static_cast<int>(sizeof(T)) && !(is_float && (sz == 4 || sz == 8)))
    KALDI_ERR << "ReadBasicType: expected size byte " << sizeof(T) << ", saw " << sz << " at file position " << is.tellg();
  if (is_float && sz != static_cast<int>(sizeof(T))) {       // float <-> double on disk
    if (sz == 4) { float f; is.read(reinterpret_cast<char*>(&f), 4); *v = static_cast<T>(f); }
    else { double d; is.read(reinterpret_cast<char*>(&d), 8); *v = static_cast<T>(d); }
  } else {
    is.read(reinterpret_cast<char*>(v), sizeof(T));
  }
  if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType, file position is " << is.tellg();
}

void WriteBasicType(std::ostream& os, bool binary, int32 v) { if (binary) WriteBin(os, v); else os << v << " "; }
void WriteBasicType(std::ostream& os, bool binary, float v) { if (binary) WriteBin(os, v); else os << v << " "; }
void WriteBasicType(std::ostream& os, bool binary, double v) { if (binary) WriteBin(os, v); else os << v << " "; }
void WriteBasicType(std::ostream& os, bool binary, bool v) { os << (v ? "T" : "F"); if (!binary) os << " "; }
void ReadBasicType(std::istream& is, bool binary, int32* v) {
  if (binary) ReadBin(is, v, false); else { is >> *v; if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType (int), next char is " << static_cast<char>(is.peek()); }
}
void ReadBasicType(std::istream& is, bool binary, float* v) {
  if (binary) ReadBin(is, v, true); else { is >> *v; if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType (float)"; }
}
void ReadBasicType(std::istream& is, bool binary, double* v) {
  if (binary) ReadBin(is, v, true); else { is >> *v; if (is.fail()) KALDI_ERR << "Read failure in ReadBasicType (double)"; }
}
void ReadBasicType(std::istream& is, bool binary, bool* v) {
  if (!binary) is >> std::ws;
  const int c = is.get();
  if (c == 'T') *v = true; else if (c == 'F') *v = false; else KALDI_ERR << "Read failure in ReadBasicType<bool>: saw " << static_cast<char>(c);
}

void WriteIntegerVector(std::ostream& os, bool binary, const std::vector<int32>& v) {
  if (binary) {
    os.put(static_cast<char>(sizeof(int32)));
    const int32 n = static_cast<int32>(v.size());
    os.write(reinterpret_cast<const char*>(&n), sizeof(n));
    if (n != 0) os.write(reinterpret_cast<const char*>(v.data()), sizeof(int32) * n);
  } else {
    os << "[ ";
    for (int32 x : v) os << x << " ";
    os << "]\n";
  }
  if (os.fail()) KALDI_ERR << "Write failure in WriteIntegerVector.";
}

void ReadIntegerVector(std::istream& is, bool binary, std::vector<int32>* v) {
  v->clear();
  if (binary) {
    const int sz = is.peek();
    if (sz != static_cast<int>(sizeof(int32))) KALDI_ERR << "ReadIntegerVector: expected size byte 4, saw " << sz;
    is.get();
    int32 n = 0;
    is.read(reinterpret_cast<char*>(&n), sizeof(n));
    if (is.fail() || n < 0) KALDI_ERR << "ReadIntegerVector: bad length";
    v->resize(n);
    if (n > 0) is.read(reinterpret_cast<char*>(v->data()), sizeof(int32) * n);
  } else {
    is >> std::ws;
    if (is.peek() != static_cast<int>('[')) KALDI_ERR << "ReadIntegerVector: expected to see [, saw " << static_cast<char>(is.peek());
    is.get();
    is >> std::ws;
    while (is.peek() != static_cast<int>(']')) {
      int32 x;
      is >> x >> std::ws;
      if (is.fail()) KALDI_ERR << "ReadIntegerVector: failed to read an integer";
      v->push_back(x);
    }
    is.get();
  }
  if (is.fail()) KALDI_ERR << "ReadIntegerVector: read failure at file position " << is.tellg();
}

void Input::Open(const std::string& name, bool* binary) {
  if (name == "-" || name.empty()) {
    is_ = &std::cin;
  } else {
    file_.reset(new std::ifstream(name.c_str(), std::ios_base::in | std::ios_base::binary));
    if (!file_->is_open()) KALDI_ERR << "Error opening input stream " << name;
    is_ = file_.get();
  }
  bool bin = false;
  if (is_->peek() == '\0') {
    is_->get();
    if (is_->peek() != 'B') KALDI_ERR << "Bad binary header in " << name;
    is_->get();
    bin = true;
  }
  if (binary != nullptr) *binary = bin;
  else if (bin) KALDI_ERR << "Binary header found in text-mode input " << name;
}
void Input::OpenTextMode(const std::string& name) {
  if (name == "-" || name.empty()) { is_ = &std::cin; return; }
  file_.reset(new std::ifstream(name.c_str(), std::ios_base::in));
  if (!file_->is_open()) KALDI_ERR << "Error opening input stream " << name;
  is_ = file_.get();
}

Output::Output(const std::string& name, bool binary, bool write_header) : os_(nullptr), name_(name) {
  if (name == "-" || name.empty()) {
    os_ = &std::cout;
  } else {
    file_.reset(new std::ofstream(name.c_str(), std::ios_base::out | std::ios_base::binary));
    if (!file_->is_open()) KALDI_ERR << "Error opening output stream " << name;
    os_ = file_.get();
  }
  if (write_header && binary) { os_->put('\0'); os_->put('B'); }
}
void Output::Close() {
  if (os_ == nullptr) return;
  os_->flush();
  if (os_->fail()) { os_ = nullptr; KALDI_ERR << "Error closing output stream " << name_; }
  file_.reset();
  os_ = nullptr;
}

bool ConvertStringToInteger(const std::string& s, int32* out) {
  char* end = nullptr;
  const long v = strtol(s.c_str(), &end, 10);
  if (end == s.c_str()) return false;
  while (*end && isspace(*end)) ++end;
  if (*end != '\0') return false;
  *out = static_cast<int32>(v);
  return true;
}
void SplitStringToVector(const std::string& full, const char* delim, bool omit_empty, std::vector<std::string>* out) {
  out->clear();
  size_t start = 0;
  while (true) {
    const size_t found = full.find_first_of(delim, start);
    const std::string piece = full.substr(start, found == std::string::npos ? std::string::npos : found - start);
    if (!omit_empty || !piece.empty()) out->push_back(piece);
    if (found == std::string::npos) break;
    start = found + 1;
  }
}
bool SplitStringToIntegers(const std::string& full, const char* delim, bool omit_empty, std::vector<int32>* out) {
  std::vector<std::string> parts;
  SplitStringToVector(full, delim, omit_empty, &parts);
  out->clear();
  for (const auto& p : parts) {
    int32 v;
    if (!ConvertStringToInteger(p, &v)) return false;
    out->push_back(v);
  }
  return true;
}

}  // namespace kaldi
