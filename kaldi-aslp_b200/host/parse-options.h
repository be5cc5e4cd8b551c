// parse-options.h -- the slice of Kaldi's ParseOptions the trainer CLIs use (src/util/parse-options.{h,cc}):
// --name=value, --name value-less booleans (--flag / --flag=true|false), --config=file, --verbose=N, --help,
// positional arguments, PrintUsage().  '-' and '_' in option names are interchangeable, which is what makes
// both `--right-splice` and `--right_splice` work.
#ifndef ASLP_HOST_PARSE_OPTIONS_H_
#define ASLP_HOST_PARSE_OPTIONS_H_
#include <map>
#include "base.h"

namespace kaldi {

class OptionsItf {
 public:
  virtual void Register(const std::string& name, bool* ptr, const std::string& doc) = 0;
  virtual void Register(const std::string& name, int32* ptr, const std::string& doc) = 0;
  virtual void Register(const std::string& name, float* ptr, const std::string& doc) = 0;
  virtual void Register(const std::string& name, double* ptr, const std::string& doc) = 0;
  virtual void Register(const std::string& name, std::string* ptr, const std::string& doc) = 0;
  virtual ~OptionsItf() {}
};

class ParseOptions : public OptionsItf {
 public:
  explicit ParseOptions(const char* usage) : usage_(usage), help_(false), config_(""), verbose_(0) {
    Register("config", &config_, "Configuration file to read (this option may be repeated)");
    Register("help", &help_, "Print out usage message");
    Register("verbose", &verbose_, "Verbose level (higher->more logging)");
  }
  void Register(const std::string& n, bool* p, const std::string& d) { Add(n, 'b', p, d); }
  void Register(const std::string& n, int32* p, const std::string& d) { Add(n, 'i', p, d); }
  void Register(const std::string& n, float* p, const std::string& d) { Add(n, 'f', p, d); }
  void Register(const std::string& n, double* p, const std::string& d) { Add(n, 'd', p, d); }
  void Register(const std::string& n, std::string* p, const std::string& d) { Add(n, 's', p, d); }
  int Read(int argc, const char* const* argv);
  int NumArgs() const { return static_cast<int>(args_.size()); }
  std::string GetArg(int i) const;       // 1-based like Kaldi
  std::string GetOptArg(int i) const { return i <= NumArgs() ? GetArg(i) : std::string(); }
  void PrintUsage(bool print_command_line = false);
 private:
  struct Opt { char type; void* ptr; std::string doc; std::string name; };
  void Add(const std::string& name, char type, void* ptr, const std::string& doc);
  static std::string Norm(const std::string& s);
  void Set(const std::string& key, const std::string& value, bool has_value);
  void ReadConfigFile(const std::string& file);
  std::map<std::string, Opt> opts_;
  std::vector<std::string> order_;
  std::vector<std::string> args_;
  const char* usage_;
  bool help_;
  std::string config_;
  int32 verbose_;
  std::string cmdline_;
};

}  // namespace kaldi
#endif
