#include "matrix.h"
#include "cu-workspace.h"
#include <cstring>

namespace kaldi {

static aslp_stream_t g_stream = nullptr;       // compute stream
static aslp_stream_t g_side = nullptr;         // side stream (lazily created)
static thread_local aslp_stream_t t_current = nullptr;   // CuStreamScope / helper-thread override of what CuStream() hands out
static thread_local bool t_helper = false;                // a thread that went through CuThreadAttach()
static int g_device = 0;                                  // the device the compute stream was created on
static bool g_stream_made = false;
static bool g_side_pending = false;
static void* g_ev_fork = nullptr;
static void* g_ev_join = nullptr;

void CuSelectDevice(int dev) {
  if (g_stream_made) KALDI_ERR << "CuSelectDevice must be called before the first device operation";
  ASLP_OK(aslp_set_device(dev));
}
aslp_stream_t CuStream() {
  if (!g_stream_made) {
    int n = 0;
    if (aslp_device_count(&n) != 0 || n <= 0)
      KALDI_ERR << "No CUDA device: this build has no CPU path (the reference's --use-gpu=no branch is the oracle, not the product)";
    ASLP_OK(aslp_get_device(&g_device));
    ASLP_OK(aslp_stream_create(&g_stream));
    g_stream_made = true;
  }
  return t_current != nullptr ? t_current : g_stream;
}
void CuThreadAttach() {
  KALDI_ASSERT(!t_helper && g_stream_made);
  ASLP_OK(aslp_set_device(g_device));                     // the current device is per-thread state
  ASLP_OK(aslp_stream_create(&t_current));
  t_helper = true;
}
void CuThreadUseDevice() { if (g_stream_made) ASLP_OK(aslp_set_device(g_device)); }
void CuWorkspaceRelease(aslp_stream_t st);
void CuThreadDetach() {
  if (!t_helper) return;
  aslp_stream_sync(t_current);
  CuWorkspaceRelease(t_current);                           // the stream handle is about to die: drop what was keyed by it
  aslp_scratch_release(t_current);
  aslp_stream_destroy(t_current);
  t_current = nullptr;
  t_helper = false;
}
aslp_stream_t CuSideStream() {
  CuStream();
  if (g_side == nullptr) ASLP_OK(aslp_stream_create(&g_side));
  return g_side;
}
bool CuOnComputeStream() { return t_current == nullptr && !t_helper; }
bool CuAsyncEnabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ASLP_ASYNC_WGRAD"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
void CuFork() {
  CuSideStream();
  ASLP_OK(aslp_event_record(g_stream, &g_ev_fork));
  ASLP_OK(aslp_stream_wait_event(g_side, g_ev_fork));
  g_side_pending = true;
}
void CuJoin() {
  if (!g_side_pending) return;
  ASLP_OK(aslp_event_record(g_side, &g_ev_join));
  ASLP_OK(aslp_stream_wait_event(g_stream, g_ev_join));
  g_side_pending = false;
}
// ------------------------------------------------------------------ step replay
bool CuStepGraph::Enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("ASLP_STEP_GRAPH"); on = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return on == 1;
}
CuStepGraph::CuStepGraph() : recording_(false), epoch_(0), clock_(0), replays_(0), recordings_(0) {}
CuStepGraph::~CuStepGraph() { DropAll(); }
void CuStepGraph::DropAll() {
  for (auto& kv : cache_) if (kv.second.exec != nullptr) aslp_graph_destroy(kv.second.exec);
  cache_.clear();
}
bool CuStepGraph::Begin(const std::string& key) {
  KALDI_ASSERT(!recording_);
  if (!Enabled() || t_helper) return true;
  const unsigned long long epoch = aslp_alloc_epoch();
  if (epoch != epoch_) {                                  // some device allocation went away: every recording may hold a dead pointer
    for (auto& kv : cache_) if (kv.second.exec != nullptr) { aslp_graph_destroy(kv.second.exec); kv.second.exec = nullptr; kv.second.seen = 0; }
    epoch_ = epoch;
  }
  auto it = cache_.find(key);
  if (it == cache_.end()) {
    if (static_cast<int>(cache_.size()) >= kMaxGraphs) {  // least recently used out
      auto victim = cache_.begin();
      for (auto jt = cache_.begin(); jt != cache_.end(); ++jt) if (jt->second.stamp < victim->second.stamp) victim = jt;
      if (victim->second.exec != nullptr) aslp_graph_destroy(victim->second.exec);
      cache_.erase(victim);
    }
    Entry e; e.exec = nullptr; e.kernels = 0; e.seen = 0; e.bad = false; e.stamp = 0;
    it = cache_.insert(std::make_pair(key, e)).first;
  }
  Entry& e = it->second;
  e.stamp = ++clock_;
  if (e.exec != nullptr) {
    ASLP_OK(aslp_graph_launch(e.exec, g_stream));
    aslp_count_launches(e.kernels);
    ++replays_;
    return false;
  }
  if (e.bad || e.seen < 2) { ++e.seen; return true; }     // warm-up passes run as they are
  CuStream();
  if (aslp_graph_begin(g_stream) != 0) { e.bad = true; return true; }
  recording_ = true;
  recording_key_ = key;
  return true;
}
bool CuStepGraph::End() {
  if (!recording_) return true;
  recording_ = false;
  Entry& e = cache_[recording_key_];
  void* exec = nullptr;
  int kernels = 0;
  if (aslp_graph_end(g_stream, &exec, &kernels) != 0 || exec == nullptr) {
    g_side_pending = false;                                // a fork recorded into the dead capture never happened
    e.bad = true;
    KALDI_WARN << "step recording failed (" << aslp_last_error() << "); this step shape keeps running unrecorded";
    return false;
  }
  if (aslp_alloc_epoch() != epoch_) {                      // something was freed while recording: the graph may point at it
    aslp_graph_destroy(exec);
    epoch_ = aslp_alloc_epoch();
    e.seen = 0;
    return false;
  }
  e.exec = exec; e.kernels = kernels;
  ++recordings_;
  ASLP_OK(aslp_graph_launch(e.exec, g_stream));            // recording does not execute: run the step now
  return true;
}

// GEMMs issued inside a side-stream scope can run on a capped grid (ASLP_SIDE_GEMM_CTAS = n; default 0 = one CTA per SM): they are
// the weight-gradient products that run under the next layer's persistent backward recurrence (80 co-resident CTAs).  Measured on
// one box (profiles/r02_ab_side_ctas.jsonl): 17.91 ms per cfg3 step uncapped, 17.95 at 68, 17.98 at 60, 18.28 at 40 -- the cap
// does not help (the recurrence gets its SMs either way), so it stays off.
static int SideGemmCtas() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("ASLP_SIDE_GEMM_CTAS"); v = e != nullptr ? atoi(e) : 0; }
  return v;
}
CuStreamScope::CuStreamScope(aslp_stream_t s) {
  CuStream(); saved_ = t_current; t_current = s;
  if (s == g_side && g_side != nullptr) aslp_gemm_set_cta_limit(SideGemmCtas());
}
CuStreamScope::~CuStreamScope() {
  if (t_current == g_side && g_side != nullptr) aslp_gemm_set_cta_limit(0);
  t_current = saved_;
}
void CuSync() {
  CuStream();
  if (t_helper) { ASLP_OK(aslp_stream_sync(t_current)); return; }     // a helper thread owns nothing but its stream
  if (g_side != nullptr) ASLP_OK(aslp_stream_sync(g_side));
  g_side_pending = false;
  ASLP_OK(aslp_stream_sync(g_stream));
}

// ------------------------------------------------------------------ host containers
PinnedMatrix::~PinnedMatrix() { if (d_ != nullptr) aslp_free_host(d_); }
void PinnedMatrix::Resize(int32 rows, int32 cols, MatrixResizeType t) {
  const size_t n = static_cast<size_t>(rows) * cols;
  if (n > cap_) {
    if (d_ != nullptr) { ASLP_OK(aslp_free_host(d_)); d_ = nullptr; }
    const size_t cap = n + n / 4 + 1024;
    void* p = nullptr;
    ASLP_OK(aslp_malloc_host(&p, cap * sizeof(float)));
    d_ = static_cast<float*>(p);
    cap_ = cap;
  }
  r_ = rows; c_ = cols;
  if (t == kSetZero && n > 0) std::memset(d_, 0, n * sizeof(float));
}
void PinnedMatrix::AppendRows(const float* src, int32 rows, int32 cols) {
  if (r_ == 0) c_ = cols;
  KALDI_ASSERT(cols == c_);
  const size_t have = static_cast<size_t>(r_) * c_, add = static_cast<size_t>(rows) * cols;
  if (have + add > cap_) {
    const size_t cap = (have + add) + (have + add) / 2 + 1024;
    void* p = nullptr;
    ASLP_OK(aslp_malloc_host(&p, cap * sizeof(float)));
    if (have > 0) std::memcpy(p, d_, have * sizeof(float));
    if (d_ != nullptr) ASLP_OK(aslp_free_host(d_));
    d_ = static_cast<float*>(p);
    cap_ = cap;
  }
  if (add > 0) std::memcpy(d_ + have, src, add * sizeof(float));
  r_ += rows;
}
template <typename Real> static const char* MatTok() { return sizeof(Real) == 4 ? "FM" : "DM"; }
template <typename Real> static const char* VecTok() { return sizeof(Real) == 4 ? "FV" : "DV"; }

template <typename Real>
void VectorBase<Real>::Write(std::ostream& os, bool binary) const {
  if (binary) {
    WriteToken(os, binary, VecTok<Real>());
    WriteBasicType(os, binary, static_cast<int32>(Dim()));
    os.write(reinterpret_cast<const char*>(data_), sizeof(Real) * dim_);
  } else {
    os << " [ ";
    for (int32 i = 0; i < dim_; ++i) os << data_[i] << " ";
    os << "]\n";
  }
  if (!os.good()) KALDI_ERR << "Failed to write vector to stream";
}

template <typename Real>
void Vector<Real>::Read(std::istream& is, bool binary) {
  if (binary) {
    std::string tok;
    ReadToken(is, binary, &tok);
    int32 dim = 0;
    ReadBasicType(is, binary, &dim);
    if (dim < 0) KALDI_ERR << "Vector::Read, negative dimension";
    d_.resize(dim);
    if (tok == VecTok<Real>()) {
      is.read(reinterpret_cast<char*>(d_.data()), sizeof(Real) * dim);
    } else if (tok == "FV") {
      std::vector<float> tmp(dim); is.read(reinterpret_cast<char*>(tmp.data()), 4 * dim);
      for (int32 i = 0; i < dim; ++i) d_[i] = static_cast<Real>(tmp[i]);
    } else if (tok == "DV") {
      std::vector<double> tmp(dim); is.read(reinterpret_cast<char*>(tmp.data()), 8 * dim);
      for (int32 i = 0; i < dim; ++i) d_[i] = static_cast<Real>(tmp[i]);
    } else {
      KALDI_ERR << "Vector::Read, expected token FV or DV, got " << tok;
    }
  } else {
    std::string s;
    is >> s;
    if (s != "[") KALDI_ERR << "Vector::Read, expected \"[\" but got " << s;
    d_.clear();
    while (true) {
      is >> std::ws;
      const int c = is.peek();
      if (c == ']') { is.get(); break; }
      if (c == EOF) KALDI_ERR << "Vector::Read, EOF while reading vector";
      std::string item;
      is >> item;
      d_.push_back(static_cast<Real>(strtod(item.c_str(), nullptr)));   // handles nan / inf spellings of the stream writer
    }
    if (is.peek() == '\r') is.get();
    if (is.peek() == '\n') is.get();
  }
  Sync();
  if (is.fail()) KALDI_ERR << "Failed to read vector from stream";
}

template <typename Real>
void MatrixBase<Real>::Write(std::ostream& os, bool binary) const {
  if (binary) {
    WriteToken(os, binary, MatTok<Real>());
    WriteBasicType(os, binary, r_);
    WriteBasicType(os, binary, c_);
    for (int32 i = 0; i < r_; ++i) os.write(reinterpret_cast<const char*>(RowData(i)), sizeof(Real) * c_);
  } else if (c_ == 0) {
    os << " [ ]\n";
  } else {
    os << " [";
    for (int32 i = 0; i < r_; ++i) {
      os << "\n  ";
      for (int32 j = 0; j < c_; ++j) os << (*this)(i, j) << " ";
    }
    os << "]\n";
  }
  if (!os.good()) KALDI_ERR << "Failed to write matrix to stream";
}

template <typename Real>
void Matrix<Real>::Read(std::istream& is, bool binary) {
  if (binary) {
    std::string tok;
    ReadToken(is, binary, &tok);
    if (tok == "CM" || tok == "CM2") {
      // CompressedMatrix (src/matrix/compressed-matrix.{h,cc}:128-143, 438-483, 486-530): what copy-feats --compress=true writes.
      // Header after the token: min_value, range (float), num_rows, num_cols (int32).  CM: per column four uint16 percentiles,
      // then one byte per element COLUMN by column, piecewise linear between the percentiles; CM2: uint16 per element, row-major.
      struct { float min_value, range; int32 num_rows, num_cols; } h;
      is.read(reinterpret_cast<char*>(&h), sizeof(h));
      if (is.fail()) KALDI_ERR << "Failed to read header";
      if (h.num_rows < 0 || h.num_cols < 0) KALDI_ERR << "Matrix::Read, negative dimension in a compressed matrix";
      Resize(h.num_cols == 0 ? 0 : h.num_rows, h.num_cols, kUndefined);
      if (h.num_cols == 0) return;
      auto u16 = [&](uint16_t v) { return h.min_value + h.range * 1.52590218966964e-05F * v; };
      if (tok == "CM") {
        std::vector<uint16_t> pc(static_cast<size_t>(h.num_cols) * 4);
        is.read(reinterpret_cast<char*>(pc.data()), pc.size() * 2);
        std::vector<unsigned char> col(h.num_rows);
        for (int32 j = 0; j < h.num_cols; ++j) {
          const float p0 = u16(pc[4 * j]), p25 = u16(pc[4 * j + 1]), p75 = u16(pc[4 * j + 2]), p100 = u16(pc[4 * j + 3]);
          is.read(reinterpret_cast<char*>(col.data()), h.num_rows);
          for (int32 i = 0; i < h.num_rows; ++i) {
            const unsigned char v = col[i];
            float f;
            if (v <= 64) f = p0 + (p25 - p0) * v * (1 / 64.0);
            else if (v <= 192) f = p25 + (p75 - p25) * (v - 64) * (1 / 128.0);
            else f = p75 + (p100 - p75) * (v - 192) * (1 / 63.0);
            (*this)(i, j) = static_cast<Real>(f);
          }
        }
      } else {
        std::vector<uint16_t> row(h.num_cols);
        for (int32 i = 0; i < h.num_rows; ++i) {
          is.read(reinterpret_cast<char*>(row.data()), static_cast<size_t>(h.num_cols) * 2);
          for (int32 j = 0; j < h.num_cols; ++j) (*this)(i, j) = static_cast<Real>(u16(row[j]));
        }
      }
      if (is.fail()) KALDI_ERR << "Failed to read data.";
      return;
    }
    int32 rows = 0, cols = 0;
    ReadBasicType(is, binary, &rows);
    ReadBasicType(is, binary, &cols);
    if (rows < 0 || cols < 0) KALDI_ERR << "Matrix::Read, negative dimension";
    Resize(rows, cols, kUndefined);
    const size_t n = static_cast<size_t>(rows) * cols;
    if (tok == MatTok<Real>()) {
      is.read(reinterpret_cast<char*>(d_.data()), sizeof(Real) * n);
    } else if (tok == "FM") {
      std::vector<float> tmp(n); is.read(reinterpret_cast<char*>(tmp.data()), 4 * n);
      for (size_t i = 0; i < n; ++i) d_[i] = static_cast<Real>(tmp[i]);
    } else if (tok == "DM") {
      std::vector<double> tmp(n); is.read(reinterpret_cast<char*>(tmp.data()), 8 * n);
      for (size_t i = 0; i < n; ++i) d_[i] = static_cast<Real>(tmp[i]);
    } else {
      KALDI_ERR << "Matrix::Read, expected token FM, DM, CM or CM2, got " << tok;
    }
  } else {
    std::string s;
    is >> s;
    if (s != "[") KALDI_ERR << "Matrix::Read, expected \"[\" but got " << s;
    std::vector<std::vector<Real>> rows;
    std::vector<Real> cur;
    while (true) {
      const int c = is.peek();
      if (c == EOF) KALDI_ERR << "Matrix::Read, EOF while reading matrix";
      if (c == ']') { is.get(); if (!cur.empty()) rows.push_back(cur); break; }
      if (c == '\n' || c == ';') { is.get(); if (!cur.empty()) { rows.push_back(cur); cur.clear(); } continue; }
      if (isspace(c)) { is.get(); continue; }
      std::string item;
      is >> item;
      if (!item.empty() && item.back() == ']') {          // "1.0]" glued
        item.pop_back();
        if (!item.empty()) cur.push_back(static_cast<Real>(strtod(item.c_str(), nullptr)));
        if (!cur.empty()) rows.push_back(cur);
        break;
      }
      cur.push_back(static_cast<Real>(strtod(item.c_str(), nullptr)));
    }
    if (is.peek() == '\r') is.get();
    if (is.peek() == '\n') is.get();
    const int32 nr = static_cast<int32>(rows.size()), nc = nr ? static_cast<int32>(rows[0].size()) : 0;
    Resize(nr, nc, kUndefined);
    for (int32 i = 0; i < nr; ++i) {
      if (static_cast<int32>(rows[i].size()) != nc) KALDI_ERR << "Matrix::Read, rows of different length";
      for (int32 j = 0; j < nc; ++j) (*this)(i, j) = rows[i][j];
    }
  }
  if (is.fail()) KALDI_ERR << "Failed to read matrix from stream";
}

template class VectorBase<float>;
template class VectorBase<double>;
template class MatrixBase<float>;
template class MatrixBase<double>;
template class Vector<float>;
template class Vector<double>;
template class Matrix<float>;
template class Matrix<double>;

// ------------------------------------------------------------------ device matrix
CuSubMatrix<float> CuMatrixBase<float>::RowRange(int32 r0, int32 n) const {
  KALDI_ASSERT(r0 >= 0 && n >= 0 && r0 + n <= rows_);
  return CuSubMatrix<float>(data_ + static_cast<size_t>(r0) * stride_, n, cols_, stride_);
}
CuSubMatrix<float> CuMatrixBase<float>::ColRange(int32 c0, int32 n) const {
  KALDI_ASSERT(c0 >= 0 && n >= 0 && c0 + n <= cols_);
  return CuSubMatrix<float>(data_ + c0, rows_, n, stride_);
}
CuSubMatrix<float> CuMatrixBase<float>::Range(int32 r0, int32 nr, int32 c0, int32 nc) const {
  KALDI_ASSERT(r0 >= 0 && nr >= 0 && r0 + nr <= rows_ && c0 >= 0 && nc >= 0 && c0 + nc <= cols_);
  return CuSubMatrix<float>(data_ + static_cast<size_t>(r0) * stride_ + c0, nr, nc, stride_);
}
void CuMatrixBase<float>::SetZero() {
  if (rows_ == 0 || cols_ == 0) return;
  if (cols_ == stride_) { ASLP_OK(aslp_memset(CuStream(), data_, 0, sizeof(float) * static_cast<size_t>(rows_) * stride_)); return; }
  ASLP_OK(aslp_memset2d(CuStream(), data_, sizeof(float) * stride_, 0, sizeof(float) * cols_, rows_));
}
void CuMatrixBase<float>::CopyFromMat(const CuMatrixBase<float>& src) {
  KALDI_ASSERT(src.NumRows() == rows_ && src.NumCols() == cols_);
  ASLP_OK(aslp_memcpy2d_d2d(CuStream(), data_, sizeof(float) * stride_, src.Data(), sizeof(float) * src.Stride(), sizeof(float) * cols_, rows_));
}
void CuMatrixBase<float>::CopyFromMat(const Matrix<float>& src) {
  KALDI_ASSERT(src.NumRows() == rows_ && src.NumCols() == cols_);
  CopyFromHost(src.Data(), src.Stride());
  CuSync();   // the host matrix may be a temporary
}
void CuMatrixBase<float>::CopyFromHost(const float* src, int32 src_stride) {
  ASLP_OK(aslp_memcpy2d_h2d(CuStream(), data_, sizeof(float) * stride_, src, sizeof(float) * src_stride, sizeof(float) * cols_, rows_));
}
void CuMatrixBase<float>::CopyToMat(Matrix<float>* dst) const {
  if (dst->NumRows() != rows_ || dst->NumCols() != cols_) dst->Resize(rows_, cols_, kUndefined);
  CopyToHost(dst->Data(), dst->Stride());
  CuSync();
}
void CuMatrixBase<float>::CopyToHost(float* dst, int32 dst_stride) const {
  ASLP_OK(aslp_memcpy2d_d2h(CuStream(), dst, sizeof(float) * dst_stride, data_, sizeof(float) * stride_, sizeof(float) * cols_, rows_));
}
void CuMatrixBase<float>::AddMat(float alpha, const CuMatrixBase<float>& A) {
  KALDI_ASSERT(A.NumRows() == rows_ && A.NumCols() == cols_);
  ASLP_OK(aslp_axpby(CuStream(), data_, stride_, A.Data(), A.Stride(), rows_, cols_, alpha, 1.0f));
}
void CuMatrixBase<float>::Scale(float alpha) {
  ASLP_OK(aslp_axpby(CuStream(), data_, stride_, data_, stride_, rows_, cols_, alpha, 0.0f));
}
double CuMatrixBase<float>::Sum() const {
  static double* dev2 = nullptr;
  if (dev2 == nullptr) ASLP_OK(aslp_malloc(reinterpret_cast<void**>(&dev2), 2 * sizeof(double)));
  ASLP_OK(aslp_sum_check(CuStream(), data_, stride_, rows_, cols_, dev2));
  double h[2];
  ASLP_OK(aslp_memcpy_d2h(CuStream(), h, dev2, sizeof(h)));
  CuSync();
  return h[1] > 0 ? std::nan("") : h[0];
}

CuMatrix<float>::~CuMatrix() { if (data_ != nullptr) aslp_free(data_); }
void CuMatrix<float>::Resize(int32 rows, int32 cols, MatrixResizeType t) {
  KALDI_ASSERT(rows >= 0 && cols >= 0);
  const int32 stride = (cols + 3) / 4 * 4;
  const size_t need = static_cast<size_t>(rows) * stride;
  if (need > cap_) {
    KALDI_ASSERT(t != kCopyData);
    CuStream();                                   // make sure the device is up before the first allocation
    if (data_ != nullptr) { CuSync(); aslp_free(data_); data_ = nullptr; }
    void* p = nullptr;
    ASLP_OK(aslp_malloc(&p, sizeof(float) * (need + 4)));
    data_ = static_cast<float*>(p);
    cap_ = need;
  }
  rows_ = rows; cols_ = cols; stride_ = stride;
  if (t == kSetZero && need > 0) ASLP_OK(aslp_memset(CuStream(), data_, 0, sizeof(float) * need));
}
void CuMatrix<float>::Swap(CuMatrix<float>* o) {
  std::swap(data_, o->data_); std::swap(rows_, o->rows_); std::swap(cols_, o->cols_); std::swap(stride_, o->stride_); std::swap(cap_, o->cap_);
}
void CuMatrix<float>::Read(std::istream& is, bool binary) {
  Matrix<float> tmp;
  tmp.Read(is, binary);
  *this = tmp;
}
void CuMatrix<float>::Write(std::ostream& os, bool binary) const {
  Matrix<float> tmp;
  CopyToMat(&tmp);
  tmp.Write(os, binary);
}

// ---- forwarder vocabulary: min / max / scalar add / log through the one-pass posterior kernel (aslp_posterior_finalize)
static void MinMax(const CuMatrixBase<float>& m, float* mn, float* mx) {
  if (m.NumRows() == 0 || m.NumCols() == 0) { *mn = 0.f; *mx = 0.f; return; }
  float* stats = static_cast<float*>(CuWorkspace(8 * sizeof(float)));
  // no log, no shift, no priors: the pass rewrites every element with itself and leaves min / max of the input in stats[0..1]
  ASLP_OK(aslp_posterior_finalize(CuStream(), const_cast<float*>(m.Data()), m.Stride(), m.NumRows(), m.NumCols(), 0, 0.f, 0.f, nullptr, 0.f, stats));
  float h[5];
  ASLP_OK(aslp_memcpy_d2h(CuStream(), h, stats, sizeof(h)));
  CuSync();
  *mn = h[0]; *mx = h[1];
}
float CuMatrixBase<float>::Min() const { float a, b; MinMax(*this, &a, &b); return a; }
float CuMatrixBase<float>::Max() const { float a, b; MinMax(*this, &a, &b); return b; }
void CuMatrixBase<float>::Set(float value) {
  if (rows_ == 0 || cols_ == 0) return;
  CuVector<float> v(cols_, kUndefined);
  v.Set(value);
  ASLP_OK(aslp_add_vec_to_rows(CuStream(), data_, stride_, rows_, cols_, v.Data(), 1.0f, 0.0f));
  CuSync();                                                 // `v` is released when we return
}
void CuMatrixBase<float>::Add(float value) {
  if (rows_ == 0 || cols_ == 0) return;
  CuVector<float> v(cols_, kUndefined);
  v.Set(value);
  ASLP_OK(aslp_add_vec_to_rows(CuStream(), data_, stride_, rows_, cols_, v.Data(), 1.0f, 1.0f));
  CuSync();
}
void CuMatrixBase<float>::ApplySoftMaxPerRow(const CuMatrixBase<float>& src) {
  KALDI_ASSERT(src.NumRows() == rows_ && src.NumCols() == cols_);
  ASLP_OK(aslp_softmax_rows(CuStream(), data_, stride_, src.Data(), src.Stride(), rows_, cols_));
}
void CuMatrixBase<float>::ApplyLog() {
  if (rows_ == 0 || cols_ == 0) return;
  float* stats = static_cast<float*>(CuWorkspace(8 * sizeof(float)));
  ASLP_OK(aslp_posterior_finalize(CuStream(), data_, stride_, rows_, cols_, 1, 0.f, 0.f, nullptr, 0.f, stats));
}

// ---- CuSubVector
template <typename Real> void CuSubVector<Real>::CopyFromVec(const CuSubVector<Real>& src) {
  KALDI_ASSERT(src.Dim() == dim_);
  if (dim_ > 0) ASLP_OK(aslp_memcpy_d2d(CuStream(), data_, src.Data(), sizeof(Real) * dim_));
}
template <typename Real> void CuSubVector<Real>::CopyFromVec(const CuVector<Real>& src) {
  KALDI_ASSERT(src.Dim() == dim_);
  if (dim_ > 0) ASLP_OK(aslp_memcpy_d2d(CuStream(), data_, src.Data(), sizeof(Real) * dim_));
}
template <typename Real> void CuSubVector<Real>::CopyFromVec(const VectorBase<Real>& src) {
  KALDI_ASSERT(src.Dim() == dim_);
  if (dim_ > 0) { ASLP_OK(aslp_memcpy_h2d(CuStream(), data_, src.Data(), sizeof(Real) * dim_)); CuSync(); }
}
template <typename Real> void CuSubVector<Real>::CopyToVec(VectorBase<Real>* dst) const {
  KALDI_ASSERT(dst->Dim() == dim_);
  if (dim_ > 0) { ASLP_OK(aslp_memcpy_d2h(CuStream(), dst->Data(), data_, sizeof(Real) * dim_)); CuSync(); }
}
template <typename Real> void CuSubVector<Real>::SetZero() { if (dim_ > 0) ASLP_OK(aslp_memset(CuStream(), data_, 0, sizeof(Real) * dim_)); }
template class CuSubVector<float>;

// ------------------------------------------------------------------ device vectors
template <typename Real> CuVector<Real>::~CuVector() { if (data_ != nullptr) aslp_free(data_); }
template <typename Real>
void CuVector<Real>::Resize(int32 dim, MatrixResizeType t) {
  if (static_cast<size_t>(dim) > cap_) {
    CuStream();
    if (data_ != nullptr) { CuSync(); aslp_free(data_); data_ = nullptr; }
    void* p = nullptr;
    ASLP_OK(aslp_malloc(&p, sizeof(Real) * (dim + 4)));
    data_ = static_cast<Real*>(p);
    cap_ = dim;
    ASLP_OK(aslp_memset(CuStream(), data_, 0, sizeof(Real) * (dim + 4)));   // pad lanes read by 128-bit loads stay finite
  }
  dim_ = dim;
  if (t == kSetZero) SetZero();
}
template <typename Real> void CuVector<Real>::SetZero() { if (dim_ > 0) ASLP_OK(aslp_memset(CuStream(), data_, 0, sizeof(Real) * dim_)); }
template <typename Real> void CuVector<Real>::Set(Real v) {
  Vector<Real> h(dim_);
  for (int32 i = 0; i < dim_; ++i) h(i) = v;
  CopyFromVec(h);
}
template <typename Real> void CuVector<Real>::CopyToVec(Vector<Real>* dst) const {
  if (dst->Dim() != dim_) dst->Resize(dim_, kUndefined);
  if (dim_ > 0) ASLP_OK(aslp_memcpy_d2h(CuStream(), dst->Data(), data_, sizeof(Real) * dim_));
  CuSync();
}
template <typename Real> void CuVector<Real>::CopyFromVec(const VectorBase<Real>& src) {
  KALDI_ASSERT(src.Dim() == dim_);
  if (dim_ > 0) ASLP_OK(aslp_memcpy_h2d(CuStream(), data_, src.Data(), sizeof(Real) * dim_));
  CuSync();
}
template <typename Real> CuVector<Real>& CuVector<Real>::operator=(const CuVector& o) {
  if (this == &o) return *this;
  Resize(o.dim_, kUndefined);
  if (dim_ > 0) ASLP_OK(aslp_memcpy_d2d(CuStream(), data_, o.data_, sizeof(Real) * dim_));
  return *this;
}
template <typename Real> CuVector<Real>& CuVector<Real>::operator=(const VectorBase<Real>& o) {
  Resize(o.Dim(), kUndefined);
  CopyFromVec(o);
  return *this;
}
template <typename Real> void CuVector<Real>::Read(std::istream& is, bool binary) { Vector<Real> t; t.Read(is, binary); *this = t; }
template <typename Real> void CuVector<Real>::Write(std::ostream& os, bool binary) const { Vector<Real> t; CopyToVec(&t); t.Write(os, binary); }
template class CuVector<float>;
template class CuVector<double>;

CuArrayInt::~CuArrayInt() { if (data_ != nullptr) aslp_free(data_); }
CuArrayInt& CuArrayInt::operator=(const std::vector<int32>& v) {
  std::vector<int32> copy(v);   // v may alias host_
  if (copy.size() > cap_) {
    CuStream();
    if (data_ != nullptr) { CuSync(); aslp_free(data_); data_ = nullptr; }
    void* p = nullptr;
    ASLP_OK(aslp_malloc(&p, sizeof(int32) * (copy.size() + 4)));
    data_ = static_cast<int32*>(p);
    cap_ = copy.size();
  }
  host_.swap(copy);
  dim_ = static_cast<int32>(host_.size());
  if (dim_ > 0) { ASLP_OK(aslp_memcpy_h2d(CuStream(), data_, host_.data(), sizeof(int32) * dim_)); CuSync(); }
  return *this;
}

// MomentStatistics (src/aslp-nnet/nnet-utils.h): mean / stddev / skewness / kurtosis summary used by Info()
static std::string Moments(const std::vector<float>& v) {
  if (v.empty()) return " ( empty )";
  double mean = 0;
  for (float x : v) mean += x;
  mean /= v.size();
  double m2 = 0, m3 = 0, m4 = 0, mn = v[0], mx = v[0];
  for (float x : v) { const double d = x - mean; m2 += d * d; m3 += d * d * d; m4 += d * d * d * d; if (x < mn) mn = x; if (x > mx) mx = x; }
  m2 /= v.size(); m3 /= v.size(); m4 /= v.size();
  std::ostringstream os;
  os << " ( min " << mn << ", max " << mx << ", mean " << mean << ", variance " << m2
     << ", skewness " << (m2 > 0 ? m3 / pow(m2, 1.5) : 0.0) << ", kurtosis " << (m2 > 0 ? m4 / (m2 * m2) - 3.0 : 0.0) << " ) ";
  return os.str();
}
std::string MomentStatistics(const CuMatrixBase<float>& m) {
  Matrix<float> h;
  m.CopyToMat(&h);
  return Moments(std::vector<float>(h.Data(), h.Data() + static_cast<size_t>(h.NumRows()) * h.NumCols()));
}
std::string MomentStatistics(const CuVector<float>& v) {
  Vector<float> h;
  v.CopyToVec(&h);
  return Moments(std::vector<float>(h.Data(), h.Data() + h.Dim()));
}

}  // namespace kaldi

// ------------------------------------------------------------------ shared workspace, GEMM precision
#include "cu-workspace.h"
#include <cstring>
#include <map>
#include <mutex>
namespace kaldi {
// one workspace per stream (compute stream, side stream, a helper thread's stream)
namespace {
struct Ws { void* p = nullptr; size_t bytes = 0; };
std::map<aslp_stream_t, Ws> g_ws;
std::mutex g_ws_mu;
}
void CuWorkspaceRelease(aslp_stream_t st) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  auto it = g_ws.find(st);
  if (it != g_ws.end()) { if (it->second.p != nullptr) aslp_free(it->second.p); g_ws.erase(it); }
}
void* CuWorkspace(size_t bytes) {
  const aslp_stream_t st = CuStream();
  std::lock_guard<std::mutex> lk(g_ws_mu);
  Ws& w = g_ws[st];
  if (bytes > w.bytes) {
    if (w.p != nullptr) { ASLP_OK(aslp_stream_sync(st)); aslp_free(w.p); w.p = nullptr; }    // its users ran on this stream
    const size_t cap = bytes + bytes / 4 + (1u << 20);
    ASLP_OK(aslp_malloc(&w.p, cap));
    w.bytes = cap;
  }
  return w.p;
}
static int g_gemm_precision = -1;
int GemmPrecision() {
  if (g_gemm_precision < 0) {
    const char* e = getenv("ASLP_GEMM_PRECISION");
    g_gemm_precision = ASLP_GEMM_3XTF32;
    if (e != nullptr && strcmp(e, "tf32") == 0) g_gemm_precision = ASLP_GEMM_TF32;
    if (e != nullptr && strcmp(e, "fp32") == 0) g_gemm_precision = ASLP_GEMM_FP32;
  }
  return g_gemm_precision;
}
void SetGemmPrecision(int p) { g_gemm_precision = p; }
}  // namespace kaldi
