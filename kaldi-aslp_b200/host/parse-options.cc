#include "parse-options.h"
#include <fstream>
#include "io.h"

namespace kaldi {

std::string ParseOptions::Norm(const std::string& s) {
  std::string o(s);
  for (char& c : o) if (c == '_') c = '-';
  return o;
}
void ParseOptions::Add(const std::string& name, char type, void* ptr, const std::string& doc) {
  const std::string k = Norm(name);
  if (opts_.count(k) == 0) order_.push_back(k);
  opts_[k] = Opt{type, ptr, doc, name};
}
void ParseOptions::Set(const std::string& key, const std::string& value, bool has_value) {
  auto it = opts_.find(Norm(key));
  if (it == opts_.end()) { PrintUsage(true); KALDI_ERR << "Invalid option --" << key; }
  Opt& o = it->second;
  if (o.type == 'b') {
    bool v = true;
    if (has_value) {
      if (value == "true" || value == "TRUE" || value == "1" || value == "t" || value == "T" || value.empty()) v = true;
      else if (value == "false" || value == "FALSE" || value == "0" || value == "f" || value == "F") v = false;
      else KALDI_ERR << "Invalid format for boolean argument [expected true or false]: --" << key << "=" << value;
    }
    *static_cast<bool*>(o.ptr) = v;
    return;
  }
  if (!has_value) KALDI_ERR << "Option --" << key << " needs a value";
  char* end = nullptr;
  switch (o.type) {
    case 'i': { long v = strtol(value.c_str(), &end, 10); if (end == value.c_str() || *end) KALDI_ERR << "Invalid integer option \"" << value << "\""; *static_cast<int32*>(o.ptr) = static_cast<int32>(v); } break;
    case 'f': { double v = strtod(value.c_str(), &end); if (end == value.c_str() || *end) KALDI_ERR << "Invalid floating-point option \"" << value << "\""; *static_cast<float*>(o.ptr) = static_cast<float>(v); } break;
    case 'd': { double v = strtod(value.c_str(), &end); if (end == value.c_str() || *end) KALDI_ERR << "Invalid floating-point option \"" << value << "\""; *static_cast<double*>(o.ptr) = v; } break;
    case 's': *static_cast<std::string*>(o.ptr) = value; break;
  }
}
void ParseOptions::ReadConfigFile(const std::string& file) {
  std::ifstream is(file.c_str());
  if (!is.good()) KALDI_ERR << "Cannot open config file: " << file;
  std::string line;
  while (std::getline(is, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line.erase(h);
    size_t a = line.find_first_not_of(" \t\r"), b = line.find_last_not_of(" \t\r");
    if (a == std::string::npos) continue;
    line = line.substr(a, b - a + 1);
    if (line.compare(0, 2, "--") != 0) KALDI_ERR << "Reading config file " << file << ": line must start with --, got: " << line;
    const size_t eq = line.find('=');
    if (eq == std::string::npos) Set(line.substr(2), "", false); else Set(line.substr(2, eq - 2), line.substr(eq + 1), true);
  }
}
int ParseOptions::Read(int argc, const char* const* argv) {
  for (int i = 0; i < argc; ++i) { if (i) cmdline_ += " "; cmdline_ += argv[i]; }
  int i = 1;
  for (; i < argc; ++i) {
    const std::string a(argv[i]);
    if (a.compare(0, 2, "--") != 0) break;
    if (a == "--") { ++i; break; }
    const size_t eq = a.find('=');
    const std::string key = eq == std::string::npos ? a.substr(2) : a.substr(2, eq - 2);
    if (eq == std::string::npos) Set(key, "", false); else Set(key, a.substr(eq + 1), true);
    if (Norm(key) == "config") ReadConfigFile(config_);
    if (help_) { PrintUsage(); exit(0); }
  }
  for (; i < argc; ++i) args_.push_back(argv[i]);
  g_kaldi_verbose_level = verbose_;
  if (verbose_ > 0 || true) std::cerr << cmdline_ << std::endl;   // Kaldi echoes the command line
  return i;
}
std::string ParseOptions::GetArg(int i) const {
  if (i < 1 || i > NumArgs()) KALDI_ERR << "ParseOptions::GetArg, invalid index " << i;
  return args_[i - 1];
}
void ParseOptions::PrintUsage(bool print_command_line) {
  std::cerr << '\n' << usage_ << '\n';
  std::cerr << "Options:\n";
  for (const auto& k : order_) {
    const Opt& o = opts_[k];
    std::cerr << "  --" << o.name << " : " << o.doc << '\n';
  }
  if (print_command_line) std::cerr << "Command line was: " << cmdline_ << '\n';
}

}  // namespace kaldi
