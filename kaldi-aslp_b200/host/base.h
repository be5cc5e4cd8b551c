// base.h -- logging / error / type conventions the reference's host code relies on, re-implemented
// minimally: KALDI_ERR and KALDI_ASSERT throw std::runtime_error (src/base/kaldi-error.cc:146), every
// main() wraps in try/catch and returns -1.  Random numbers follow src/base/kaldi-math.h:147-154
// so that `aslp-nnet-init --seed` reproduces the reference CPU initialisation stream.
#ifndef ASLP_HOST_BASE_H_
#define ASLP_HOST_BASE_H_
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace kaldi {

typedef int32_t int32;
typedef int64_t int64;
typedef float BaseFloat;
typedef int32 MatrixIndexT;

extern int g_kaldi_verbose_level;
inline void SetVerboseLevel(int i) { g_kaldi_verbose_level = i; }      // src/base/kaldi-error.h:61-64
inline int GetVerboseLevel() { return g_kaldi_verbose_level; }

class MessageLogger {
 public:
  enum Kind { kLog, kWarn, kError, kAssert };
  MessageLogger(Kind kind, const char* func, const char* file, int line) : kind_(kind) {
    const char* base = file;
    for (const char* p = file; *p; ++p) if (*p == '/') base = p + 1;
    const char* tag = kind == kLog ? "LOG" : (kind == kWarn ? "WARNING" : "ERROR");
    ss_ << tag << " (" << func << "():" << base << ":" << line << ") ";
  }
  ~MessageLogger() noexcept(false) {
    if (kind_ == kError || kind_ == kAssert) throw std::runtime_error(ss_.str());
    std::cerr << ss_.str() << std::endl;
  }
  std::ostream& stream() { return ss_; }
 private:
  Kind kind_;
  std::ostringstream ss_;
};

#define KALDI_ERR ::kaldi::MessageLogger(::kaldi::MessageLogger::kError, __func__, __FILE__, __LINE__).stream()
#define KALDI_WARN ::kaldi::MessageLogger(::kaldi::MessageLogger::kWarn, __func__, __FILE__, __LINE__).stream()
#define KALDI_LOG ::kaldi::MessageLogger(::kaldi::MessageLogger::kLog, __func__, __FILE__, __LINE__).stream()
#define KALDI_VLOG(v) if ((v) <= ::kaldi::g_kaldi_verbose_level) KALDI_LOG
#define KALDI_ASSERT(cond) do { if (!(cond)) KALDI_ERR << "Assertion failed: " #cond; } while (0)
#define KALDI_ISFINITE(x) std::isfinite(x)

// status check for the C-ABI: non-zero -> KALDI_ERR with the library's text
#define ASLP_OK(call) do { int rc__ = (call); if (rc__ != 0) KALDI_ERR << #call << " failed (" << rc__ << "): " << aslp_last_error(); } while (0)

// ---- random numbers (src/base/kaldi-math.h:129-154, kaldi-math.cc:45-70) ----
inline int Rand() { return rand(); }
struct RandomState {
  RandomState() { seed = Rand() + 27437; }
  unsigned seed;
};
inline float RandUniform(RandomState* state = nullptr) {
  const int r = state ? rand_r(&state->seed) : Rand();
  return static_cast<float>((r + 1.0) / (RAND_MAX + 2.0));
}
inline float RandGauss(RandomState* state = nullptr) {
  const float u1 = RandUniform(state);     // the reference evaluates Log(RandUniform()) first ...
  const float u2 = RandUniform(state);     // ... then cosf(2 pi RandUniform())
  return static_cast<float>(sqrtf(-2 * logf(u1)) * cosf(2 * M_PI * u2));
}

class Timer {
 public:
  Timer() { Reset(); }
  void Reset() { t0_ = std::chrono::steady_clock::now(); }
  double Elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
 private:
  std::chrono::steady_clock::time_point t0_;
};

template <class T> inline std::string ToString(const T& t) { std::ostringstream os; os << t; return os.str(); }

}  // namespace kaldi
#endif
