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
bool CuAsyncEnabled();
void CuFork();
void CuJoin();
class CuStreamScope {
 public:
  explicit CuStreamScope(aslp_stream_t s);
  ~CuStreamScope();
 private:
  aslp_stream_t saved_;
};

template <typename Real>
class Vector {
 public:
  Vector() {}
  explicit Vector(int32 dim) : d_(dim, Real(0)) {}
  void Resize(int32 dim, MatrixResizeType t = kSetZero) { if (t == kSetZero) d_.assign(dim, Real(0)); else d_.resize(dim); }
  int32 Dim() const { return static_cast<int32>(d_.size()); }
  Real* Data() { return d_.data(); }
  const Real* Data() const { return d_.data(); }
  Real& operator()(int32 i) { return d_[i]; }
  Real operator()(int32 i) const { return d_[i]; }
  Real Sum() const { double s = 0; for (Real v : d_) s += v; return static_cast<Real>(s); }
  void Read(std::istream& is, bool binary);
  void Write(std::ostream& os, bool binary) const;
 private:
  std::vector<Real> d_;
};

template <typename Real>
class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(int32 rows, int32 cols) : r_(rows), c_(cols), d_(static_cast<size_t>(rows) * cols, Real(0)) {}
  void Resize(int32 rows, int32 cols, MatrixResizeType t = kSetZero) {
    r_ = rows; c_ = cols;
    if (t == kSetZero) d_.assign(static_cast<size_t>(rows) * cols, Real(0)); else d_.resize(static_cast<size_t>(rows) * cols);
  }
  int32 NumRows() const { return r_; }
  int32 NumCols() const { return c_; }
  int32 Stride() const { return c_; }
  Real* Data() { return d_.data(); }
  const Real* Data() const { return d_.data(); }
  Real* RowData(int32 r) { return d_.data() + static_cast<size_t>(r) * c_; }
  const Real* RowData(int32 r) const { return d_.data() + static_cast<size_t>(r) * c_; }
  Real& operator()(int32 r, int32 c) { return d_[static_cast<size_t>(r) * c_ + c]; }
  Real operator()(int32 r, int32 c) const { return d_[static_cast<size_t>(r) * c_ + c]; }
  void Read(std::istream& is, bool binary);
  void Write(std::ostream& os, bool binary) const;
 private:
  int32 r_, c_;
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
  CuVector& operator=(const Vector<Real>& o);
  ~CuVector();
  void Resize(int32 dim, MatrixResizeType t = kSetZero);
  int32 Dim() const { return dim_; }
  Real* Data() { return data_; }
  const Real* Data() const { return data_; }
  void SetZero();
  void Set(Real v);
  void CopyToVec(Vector<Real>* dst) const;     // synchronises
  void CopyFromVec(const Vector<Real>& src);
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
