///////////////////////////////////////////////////////////////////////////////
// simple-mpc (B200-native) — host-side mirror of the reference's C++ API for the MPC::iterate path.
//
// fwd.hpp: the small dense types that stand in for Eigen / pinocchio::SE3 in the reference's
// signatures (reference include/simple-mpc/fwd.hpp:21-51). Eigen and Pinocchio are not available in
// this build environment, so `VectorXd`, `MatrixXd`, `Vector6d`, `SE3` are our own value types with
// the few operations the API needs; the names are kept so that user code reads the same.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace simple_mpc
{

  /// Dynamic column vector of doubles (Eigen::VectorXd stand-in).
  class VectorXd : public std::vector<double>
  {
  public:
    using std::vector<double>::vector;
    VectorXd() = default;
    VectorXd(const std::vector<double> & v) : std::vector<double>(v) {}
    static VectorXd Zero(std::size_t n) { return VectorXd(n, 0.0); }
    void setZero() { std::fill(begin(), end(), 0.0); }
    VectorXd head(std::size_t n) const { return segment(0, n); }
    VectorXd tail(std::size_t n) const { return segment(size() - n, n); }
    VectorXd segment(std::size_t i, std::size_t n) const
    {
      if (i + n > size())
        throw std::runtime_error("VectorXd::segment out of range");
      return VectorXd(begin() + (long)i, begin() + (long)(i + n));
    }
    void setSegment(std::size_t i, const VectorXd & v)
    {
      if (i + v.size() > size())
        throw std::runtime_error("VectorXd::setSegment out of range");
      std::copy(v.begin(), v.end(), begin() + (long)i);
    }
  };
  using ConstVectorRef = const VectorXd &;
  using Vector6d = VectorXd; // size checked where the reference uses a fixed-size type
  using Vector7d = VectorXd;

  /// Dense row-major matrix of doubles (Eigen::MatrixXd stand-in).
  class MatrixXd
  {
  public:
    MatrixXd() = default;
    MatrixXd(std::size_t r, std::size_t c) : rows_(r), cols_(c), d_(r * c, 0.0) {}
    static MatrixXd Zero(std::size_t r, std::size_t c) { return MatrixXd(r, c); }
    static MatrixXd Identity(std::size_t n)
    {
      MatrixXd m(n, n);
      for (std::size_t i = 0; i < n; i++)
        m(i, i) = 1.0;
      return m;
    }
    static MatrixXd Diagonal(const std::vector<double> & d)
    {
      MatrixXd m(d.size(), d.size());
      for (std::size_t i = 0; i < d.size(); i++)
        m(i, i) = d[i];
      return m;
    }
    double & operator()(std::size_t i, std::size_t j) { return d_[i * cols_ + j]; }
    double operator()(std::size_t i, std::size_t j) const { return d_[i * cols_ + j]; }
    std::size_t rows() const { return rows_; }
    std::size_t cols() const { return cols_; }
    const double * data() const { return d_.data(); }
    double * data() { return d_.data(); }
    MatrixXd operator*(double s) const
    {
      MatrixXd m(*this);
      for (auto & v : m.d_)
        v *= s;
      return m;
    }
    bool operator==(const MatrixXd & o) const { return rows_ == o.rows_ && cols_ == o.cols_ && d_ == o.d_; }

  private:
    std::size_t rows_ = 0, cols_ = 0;
    std::vector<double> d_;
  };

  using Vector3d = std::array<double, 3>;

  /// Rigid transform (pinocchio::SE3 stand-in): x_parent = R x_child + p, R row-major.
  struct SE3
  {
    std::array<double, 9> R{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    Vector3d p{{0, 0, 0}};
    static SE3 Identity() { return SE3(); }
    Vector3d & translation() { return p; }
    const Vector3d & translation() const { return p; }
    std::array<double, 9> & rotation() { return R; }
    const std::array<double, 9> & rotation() const { return R; }
    SE3 operator*(const SE3 & o) const
    {
      SE3 r;
      for (int i = 0; i < 3; i++)
      {
        for (int j = 0; j < 3; j++)
          r.R[3 * i + j] = R[3 * i] * o.R[j] + R[3 * i + 1] * o.R[3 + j] + R[3 * i + 2] * o.R[6 + j];
        r.p[i] = R[3 * i] * o.p[0] + R[3 * i + 1] * o.p[1] + R[3 * i + 2] * o.p[2] + p[i];
      }
      return r;
    }
    SE3 inverse() const
    {
      SE3 r;
      for (int i = 0; i < 3; i++)
      {
        for (int j = 0; j < 3; j++)
          r.R[3 * i + j] = R[3 * j + i];
        r.p[i] = -(R[i] * p[0] + R[3 + i] * p[1] + R[6 + i] * p[2]);
      }
      return r;
    }
    /// this^-1 * o
    SE3 actInv(const SE3 & o) const { return inverse() * o; }
    Vector3d act(const Vector3d & x) const
    {
      Vector3d y;
      for (int i = 0; i < 3; i++)
        y[i] = R[3 * i] * x[0] + R[3 * i + 1] * x[1] + R[3 * i + 2] * x[2] + p[i];
      return y;
    }
  };

  inline Vector3d operator+(const Vector3d & a, const Vector3d & b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
  inline Vector3d operator-(const Vector3d & a, const Vector3d & b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
  inline Vector3d operator*(const Vector3d & a, double s) { return {{a[0] * s, a[1] * s, a[2] * s}}; }
  inline Vector3d cross(const Vector3d & a, const Vector3d & b)
  {
    return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
  }

} // namespace simple_mpc
