// matrix.h -- host Matrix/Vector (I/O containers, Kaldi "FM"/"FV"/"DM"/"DV" formats, src/matrix/kaldi-matrix.cc:1201-1330)
// and the device-resident CuMatrix/CuVector used by the Component mirror.  Unlike the reference's CuMatrix
// (src/aslp-cudamatrix/cu-matrix.h, 852 lines: every method = one kernel or a CPU fallback) this class is only
// a typed view of device memory plus the handful of elementwise helpers the Nnet executor needs; all arithmetic
// goes through the fused C-ABI calls of include/aslp_b200.h on ONE process-wide stream.  No CPU branch exists.
#ifndef ASLP_HOST_MATRIX_H_
#define ASLP_HOST_MATRIX_H_
#include "aslp_b200.h"
#include "base.h"
#include "io.h"
#include <map>
#include <string>

namespace kaldi {

typedef enum { kSetZero, kUndefined, kCopyData } MatrixResizeType;
typedef enum { kNoTrans = 0, kTrans = 1 } MatrixTransposeType;

aslp_stream_t CuStream();          // the process-wide compute stream (created on first use, device already selected)
void CuSelectDevice(int dev);      // CuDevice::SelectGpuId equivalent; must precede the first CuStream()
void CuSync();                     // waits for the compute stream AND the side stream
// Side stream for work that is off the critical path of a pass (the weight-gradient GEMMs and the update of a layer
// overlap the backward recurrence of the layer below).  CuFork(): later side-stream work starts after everything issued
// so far on the compute stream.  CuStreamScope(CuSideStream()): every device call of this library inside the scope goes to
// the side stream (CuStream() returns it, CuWorkspace() hands out the side stream's own workspace).  CuJoin(): the
// compute stream waits for the side stream; a no-op when nothing was forked.  ASLP_ASYNC_WGRAD=0 disables the overlap.
aslp_stream_t CuSideStream();
// Helper threads (the batch feeders of the trainer mains): CuThreadAttach() gives the calling thread the process's device
// and a stream of its own; from then on every device call this library makes on that thread goes to that stream (with its own
// workspace) and CuSync() waits for it alone.  CuThreadDetach() drains and destroys the stream.
void CuThreadAttach();
void CuThreadDetach();
// a helper thread that only touches page-locked memory and events (no stream of its own) still has per-thread device state: without
// this call its runtime calls land on device 0 and create a context there on every rank.  No-op before the first device operation.
void CuThreadUseDevice();
bool CuAsyncEnabled();
// true when CuStream() is the process-wide compute stream itself: no side-stream scope, not a helper thread's stream
bool CuOnComputeStream();
void CuFork();
void CuJoin();

// Static-shape step replay (include/aslp_b200.h, aslp_graph_*).  A trainer whose minibatch has the same shapes, buffers
// and hyper-parameters step after step records the device work of one step from the compute stream and replays it as ONE
// graph launch: the host cost of a step drops from one enqueue per kernel (4-14 us each -- the whole step time of the
// launch-bound configurations) to one call.
//     if (g.Begin(key)) { <enqueue the step on CuStream()>;  if (!g.End()) <enqueue it again>; }
// Begin() returns false when it replayed a recording for `key`; true when the caller has to enqueue the step: the first
// two times a key is seen (allocations settle), while recording (third time), when recording is switched off
// (ASLP_STEP_GRAPH=0) or failed for this key before.  End() returns false only when a recording was invalidated: nothing
// ran and the caller enqueues the step once more, unrecorded.  Everything between Begin and End must be device work on
// the library's streams -- host-side bookkeeping of the step belongs outside, it is not re-executed by a replay.
// Recordings hold raw device pointers: they are dropped as soon as the allocation epoch moves (any device free anywhere),
// and at most kMaxGraphs of them are kept (least recently used out first).
class CuStepGraph {
 public:
  CuStepGraph();
  ~CuStepGraph();
  bool Begin(const std::string& key);
  bool End();
  long long Replays() const { return replays_; }
  long long Recordings() const { return recordings_; }
  static bool Enabled();
 private:
  struct Entry { void* exec; int kernels; int seen; bool bad; unsigned long long stamp; };
  void DropAll();
  static const int kMaxGraphs = 32;
  std::map<std::string, Entry> cache_;
  std::string recording_key_;
  bool recording_;
  unsigned long long epoch_, clock_;
  long long replays_, recordings_;
};
class CuStreamScope {
 public:
  explicit CuStreamScope(aslp_stream_t s);
  ~CuStreamScope();
 private:
  aslp_stream_t saved_;
};

// Host containers with the reference's class split (src/matrix/kaldi-vector.h, kaldi-matrix.h): VectorBase / SubVector /
// Vector and MatrixBase / SubMatrix / Matrix, so that trainer code written against them (Matrix::Row(i).CopyFromVec(...),
// Vector<BaseFloat> v(n, kSetZero), const VectorBase<BaseFloat>& arguments) compiles unchanged.  Dense rows (stride == cols).
template <typename Real>
class VectorBase {
 public:
  int32 Dim() const { return dim_; }
  Real* Data() { return data_; }
  const Real* Data() const { return data_; }
  Real& operator()(int32 i) { return data_[i]; }
  Real operator()(int32 i) const { return data_[i]; }
  Real Sum() const { double s = 0; for (int32 i = 0; i < dim_; ++i) s += data_[i]; return static_cast<Real>(s); }
  void Set(Real v) { for (int32 i = 0; i < dim_; ++i) data_[i] = v; }
  void SetZero() { Set(Real(0)); }
  void Scale(Real a) { for (int32 i = 0; i < dim_; ++i) data_[i] *= a; }
  void Add(Real a) { for (int32 i = 0; i < dim_; ++i) data_[i] += a; }
  Real Max() const { Real m = dim_ ? data_[0] : Real(0); for (int32 i = 1; i < dim_; ++i) if (data_[i] > m) m = data_[i]; return m; }
  Real Min() const { Real m = dim_ ? data_[0] : Real(0); for (int32 i = 1; i < dim_; ++i) if (data_[i] < m) m = data_[i]; return m; }
  template <typename Other>
  void CopyFromVec(const VectorBase<Other>& v) {
    KALDI_ASSERT(v.Dim() == dim_);
    for (int32 i = 0; i < dim_; ++i) data_[i] = static_cast<Real>(v(i));
  }
  void AddVec(Real alpha, const VectorBase<Real>& v) { KALDI_ASSERT(v.Dim() == dim_); for (int32 i = 0; i < dim_; ++i) data_[i] += alpha * v(i); }
  void Write(std::ostream& os, bool binary) const;
 protected:
  VectorBase() : data_(nullptr), dim_(0) {}
  VectorBase(Real* d, int32 n) : data_(d), dim_(n) {}
  Real* data_;
  int32 dim_;
};

template <typename Real>
class SubVector : public VectorBase<Real> {
 public:
  SubVector(Real* d, int32 n) : VectorBase<Real>(d, n) {}
  SubVector(const VectorBase<Real>& v, int32 origin, int32 n) : VectorBase<Real>(const_cast<Real*>(v.Data()) + origin, n) { KALDI_ASSERT(origin + n <= v.Dim()); }
  SubVector(const SubVector& o) : VectorBase<Real>(o.data_, o.dim_) {}
};

template <typename Real>
class Vector : public VectorBase<Real> {
 public:
  Vector() {}
  explicit Vector(int32 dim, MatrixResizeType t = kSetZero) { Resize(dim, t); }
  Vector(const Vector& o) : VectorBase<Real>(), d_(o.d_) { Sync(); }
  explicit Vector(const VectorBase<Real>& o) : d_(o.Data(), o.Data() + o.Dim()) { Sync(); }
  Vector& operator=(const Vector& o) { if (this != &o) { d_ = o.d_; Sync(); } return *this; }
  Vector& operator=(const VectorBase<Real>& o) { if (this != &o) { d_.assign(o.Data(), o.Data() + o.Dim()); Sync(); } return *this; }
  void Resize(int32 dim, MatrixResizeType t = kSetZero) { if (t == kSetZero) d_.assign(dim, Real(0)); else d_.resize(dim); Sync(); }
  void Swap(Vector* o) { d_.swap(o->d_); Sync(); o->Sync(); }
  void Read(std::istream& is, bool binary);
 private:
  void Sync() { this->data_ = d_.data(); this->dim_ = static_cast<int32>(d_.size()); }
  std::vector<Real> d_;
};

template <typename Real> class SubMatrix;
template <typename Real>
class MatrixBase {
 public:
  int32 NumRows() const { return r_; }
  int32 NumCols() const { return c_; }
  int32 Stride() const { return stride_; }
  Real* Data() { return data_; }
  const Real* Data() const { return data_; }
  Real* RowData(int32 r) { return data_ + static_cast<size_t>(r) * stride_; }
  const Real* RowData(int32 r) const { return data_ + static_cast<size_t>(r) * stride_; }
  Real& operator()(int32 r, int32 c) { return data_[static_cast<size_t>(r) * stride_ + c]; }
  Real operator()(int32 r, int32 c) const { return data_[static_cast<size_t>(r) * stride_ + c]; }
  SubVector<Real> Row(int32 r) { KALDI_ASSERT(r >= 0 && r < r_); return SubVector<Real>(RowData(r), c_); }
  const SubVector<Real> Row(int32 r) const { KALDI_ASSERT(r >= 0 && r < r_); return SubVector<Real>(const_cast<Real*>(RowData(r)), c_); }
  void SetZero() { for (int32 r = 0; r < r_; ++r) for (int32 c = 0; c < c_; ++c) (*this)(r, c) = Real(0); }
  void Set(Real v) { for (int32 r = 0; r < r_; ++r) for (int32 c = 0; c < c_; ++c) (*this)(r, c) = v; }
  void Scale(Real a) { for (int32 r = 0; r < r_; ++r) for (int32 c = 0; c < c_; ++c) (*this)(r, c) *= a; }
  Real Sum() const { double s = 0; for (int32 r = 0; r < r_; ++r) for (int32 c = 0; c < c_; ++c) s += (*this)(r, c); return static_cast<Real>(s); }
  void CopyFromMat(const MatrixBase<Real>& m) {
    KALDI_ASSERT(m.NumRows() == r_ && m.NumCols() == c_);
    for (int32 r = 0; r < r_; ++r) for (int32 c = 0; c < c_; ++c) (*this)(r, c) = m(r, c);
  }
  void CopyRowFromVec(const VectorBase<Real>& v, int32 row) { Row(row).CopyFromVec(v); }
  SubMatrix<Real> RowRange(int32 r0, int32 n) const;     // kaldi-matrix.h:202: a view, rows r0 .. r0 + n - 1
  void Write(std::ostream& os, bool binary) const;
 protected:
  MatrixBase() : data_(nullptr), r_(0), c_(0), stride_(0) {}
  MatrixBase(Real* d, int32 r, int32 c, int32 s) : data_(d), r_(r), c_(c), stride_(s) {}
  Real* data_;
  int32 r_, c_, stride_;
};

template <typename Real>
class SubMatrix : public MatrixBase<Real> {
 public:
  SubMatrix(const MatrixBase<Real>& m, int32 r0, int32 nr, int32 c0, int32 nc)
      : MatrixBase<Real>(const_cast<Real*>(m.RowData(r0)) + c0, nr, nc, m.Stride()) { KALDI_ASSERT(r0 + nr <= m.NumRows() && c0 + nc <= m.NumCols()); }
};

template <typename Real>
inline SubMatrix<Real> MatrixBase<Real>::RowRange(int32 r0, int32 n) const { return SubMatrix<Real>(*this, r0, n, 0, c_); }

template <typename Real>
class Matrix : public MatrixBase<Real> {
 public:
  Matrix() {}
  Matrix(int32 rows, int32 cols, MatrixResizeType t = kSetZero) { Resize(rows, cols, t); }
  Matrix(const Matrix& o) : MatrixBase<Real>(), d_(o.d_) { Sync(o.r_, o.c_); }
  explicit Matrix(const MatrixBase<Real>& o) { Resize(o.NumRows(), o.NumCols(), kUndefined); this->CopyFromMat(o); }
  Matrix& operator=(const Matrix& o) { if (this != &o) { d_ = o.d_; Sync(o.r_, o.c_); } return *this; }
  Matrix& operator=(const MatrixBase<Real>& o) { if (this != &o) { Resize(o.NumRows(), o.NumCols(), kUndefined); this->CopyFromMat(o); } return *this; }
  void Resize(int32 rows, int32 cols, MatrixResizeType t = kSetZero) {
    if (t == kSetZero) d_.assign(static_cast<size_t>(rows) * cols, Real(0)); else d_.resize(static_cast<size_t>(rows) * cols);
    Sync(rows, cols);
  }
  void Swap(Matrix* o) { d_.swap(o->d_); const int32 r = this->r_, c = this->c_; Sync(o->r_, o->c_); o->Sync(r, c); }
  void Read(std::istream& is, bool binary);
 private:
  void Sync(int32 r, int32 c) { this->data_ = d_.data(); this->r_ = r; this->c_ = c; this->stride_ = c; }
  std::vector<Real> d_;
};

// page-locked host matrix (dense rows): the staging slots of the batch feeders, so that the H2D copy of a minibatch is a true
// asynchronous DMA (CuMatrixBase::CopyFromHost) instead of a driver-staged pageable copy
class PinnedMatrix {
 public:
  PinnedMatrix() : d_(nullptr), cap_(0), r_(0), c_(0) {}
  ~PinnedMatrix();
  PinnedMatrix(const PinnedMatrix&) = delete;
  PinnedMatrix& operator=(const PinnedMatrix&) = delete;
  void Resize(int32 rows, int32 cols, MatrixResizeType t = kSetZero);   // grows geometrically, never shrinks; contents are NOT kept
  void AppendRows(const float* src, int32 rows, int32 cols);            // keeps the contents (cols must match unless empty)
  int32 NumRows() const { return r_; }
  int32 NumCols() const { return c_; }
  int32 Stride() const { return c_; }
  float* Data() { return d_; }
  const float* Data() const { return d_; }
  float* RowData(int32 r) { return d_ + static_cast<size_t>(r) * c_; }
  const float* RowData(int32 r) const { return d_ + static_cast<size_t>(r) * c_; }
 private:
  float* d_;
  size_t cap_;
  int32 r_, c_;
};

// The device classes are templates on the element type like the reference's (src/aslp-cudamatrix/cu-matrix.h:48, cu-vector.h), so that
// code written against CuMatrix<BaseFloat> / CuSubMatrix<BaseFloat> / CuVector<BaseFloat> compiles unchanged; the path is fp32
// (BaseFloat = float, -DKALDI_DOUBLEPRECISION=0), so only the float specialisations are defined (CuVector also for double: the
// BatchNorm running sums).
template <typename Real> class CuMatrixBase;
template <typename Real> class CuSubMatrix;
template <typename Real> class CuMatrix;
template <typename Real> class CuVector;

// one row of a device matrix (CuMatrixBase::Row, src/aslp-cudamatrix/cu-matrix.h:575-590): what the forwarders' frame-skipping
// loops copy row by row (aslp-nnet-forward.cc:163-179)
template <typename Real>
class CuSubVector {
 public:
  CuSubVector(Real* d, int32 n) : data_(d), dim_(n) {}
  int32 Dim() const { return dim_; }
  Real* Data() { return data_; }
  const Real* Data() const { return data_; }
  void CopyFromVec(const CuSubVector<Real>& src);            // device -> device
  void CopyFromVec(const CuVector<Real>& src);
  void CopyFromVec(const VectorBase<Real>& src);             // host -> device
  void CopyToVec(VectorBase<Real>* dst) const;               // device -> host (synchronises)
  void SetZero();
 private:
  Real* data_;
  int32 dim_;
};

template <>
class CuMatrixBase<float> {
 public:
  float* Data() { return data_; }
  const float* Data() const { return data_; }
  int32 NumRows() const { return rows_; }
  int32 NumCols() const { return cols_; }
  int32 Stride() const { return stride_; }
  CuSubMatrix<float> RowRange(int32 r0, int32 n) const;
  CuSubMatrix<float> ColRange(int32 c0, int32 n) const;
  CuSubMatrix<float> Range(int32 r0, int32 nr, int32 c0, int32 nc) const;
  void SetZero();
  void CopyFromMat(const CuMatrixBase<float>& src);                 // device -> device
  void CopyFromMat(const Matrix<float>& src);                // host -> device
  void CopyFromHost(const float* src, int32 src_stride);     // host (pinned or pageable) -> device, async
  void CopyToMat(Matrix<float>* dst) const;                  // device -> host (synchronises)
  void CopyToHost(float* dst, int32 dst_stride) const;       // async
  void AddMat(float alpha, const CuMatrixBase<float>& A);           // this += alpha * A
  void Scale(float alpha);
  double Sum() const;                                        // synchronises
  CuSubVector<float> Row(int32 r) { KALDI_ASSERT(r >= 0 && r < rows_); return CuSubVector<float>(data_ + static_cast<size_t>(r) * stride_, cols_); }
  const CuSubVector<float> Row(int32 r) const { KALDI_ASSERT(r >= 0 && r < rows_); return CuSubVector<float>(data_ + static_cast<size_t>(r) * stride_, cols_); }
  // the forwarders' post-processing vocabulary (aslp-nnet-forward.cc:184-207); Min / Max synchronise
  float Min() const;
  float Max() const;
  void Add(float value);
  void ApplyLog();
  void Set(float value);
  void ApplySoftMaxPerRow(const CuMatrixBase<float>& src);   // cu-matrix.cc:1346-1380 (--add-softmax of the forwarders)
 protected:
  CuMatrixBase() : data_(nullptr), rows_(0), cols_(0), stride_(0) {}
  CuMatrixBase(float* d, int32 r, int32 c, int32 s) : data_(d), rows_(r), cols_(c), stride_(s) {}
  float* data_;
  int32 rows_, cols_, stride_;
};

template <>
class CuSubMatrix<float> : public CuMatrixBase<float> {
 public:
  CuSubMatrix(float* d, int32 r, int32 c, int32 s) : CuMatrixBase<float>(d, r, c, s) {}
};

template <>
class CuMatrix<float> : public CuMatrixBase<float> {
 public:
  CuMatrix() : cap_(0) {}
  CuMatrix(int32 rows, int32 cols, MatrixResizeType t = kSetZero) : cap_(0) { Resize(rows, cols, t); }
  CuMatrix(const CuMatrix& o) : CuMatrixBase<float>(), cap_(0) { *this = o; }
  explicit CuMatrix(const CuMatrixBase<float>& o) : cap_(0) { Resize(o.NumRows(), o.NumCols(), kUndefined); CopyFromMat(o); }
  explicit CuMatrix(const Matrix<float>& o) : cap_(0) { *this = o; }
  CuMatrix& operator=(const CuMatrix& o) { if (this != &o) { Resize(o.NumRows(), o.NumCols(), kUndefined); CopyFromMat(o); } return *this; }
  CuMatrix& operator=(const CuMatrixBase<float>& o) { Resize(o.NumRows(), o.NumCols(), kUndefined); CopyFromMat(o); return *this; }
  CuMatrix& operator=(const Matrix<float>& o) { Resize(o.NumRows(), o.NumCols(), kUndefined); CopyFromMat(o); return *this; }
  ~CuMatrix();
  // rows are 16-byte aligned (stride = cols rounded up to 4 floats), like the pitched allocation of the reference
  void Resize(int32 rows, int32 cols, MatrixResizeType t = kSetZero);
  void Swap(CuMatrix* o);
  void Read(std::istream& is, bool binary);
  void Write(std::ostream& os, bool binary) const;
 private:
  size_t cap_;    // allocated floats
};

template <typename Real>
class CuVector {
 public:
  CuVector() : data_(nullptr), dim_(0), cap_(0) {}
  explicit CuVector(int32 dim, MatrixResizeType t = kSetZero) : data_(nullptr), dim_(0), cap_(0) { Resize(dim, t); }
  CuVector(const CuVector& o) : data_(nullptr), dim_(0), cap_(0) { *this = o; }
  CuVector& operator=(const CuVector& o);
  CuVector& operator=(const VectorBase<Real>& o);
  ~CuVector();
  void Resize(int32 dim, MatrixResizeType t = kSetZero);
  int32 Dim() const { return dim_; }
  Real* Data() { return data_; }
  const Real* Data() const { return data_; }
  void SetZero();
  void Set(Real v);
  void CopyToVec(Vector<Real>* dst) const;     // synchronises
  void CopyFromVec(const VectorBase<Real>& src);
  void Read(std::istream& is, bool binary);
  void Write(std::ostream& os, bool binary) const;
 private:
  Real* data_;
  int32 dim_;
  size_t cap_;
};

// device int32 array (CuArray<int32>)
class CuArrayInt {
 public:
  CuArrayInt() : data_(nullptr), dim_(0), cap_(0) {}
  CuArrayInt(const CuArrayInt& o) : data_(nullptr), dim_(0), cap_(0) { *this = o.host_; }
  CuArrayInt& operator=(const CuArrayInt& o) { return *this = o.host_; }
  CuArrayInt& operator=(const std::vector<int32>& v);
  ~CuArrayInt();
  int32 Dim() const { return dim_; }
  const int32* Data() const { return data_; }
  const std::vector<int32>& Host() const { return host_; }
 private:
  int32* data_;
  int32 dim_;
  size_t cap_;
  std::vector<int32> host_;
};

std::string MomentStatistics(const CuMatrixBase<float>& m);
std::string MomentStatistics(const CuVector<float>& v);

}  // namespace kaldi
#endif
