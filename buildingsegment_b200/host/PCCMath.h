// PCCMath.h -- Vec3<T> / Box3<T> with the operator set the segmentation path uses.
// Mirrors the interface of the reference's tmc3/PCCMath.h:54-550 (same names, same arithmetic:
// component-wise +,-,*,/ ; operator* of two vectors is the DOT product, evaluated left to right
// as a0*b0 + a1*b1 + a2*b2 (PCCMath.h:333-338); operator/= converts the divisor type first,
// PCCMath.h:227-235).  Written from scratch; the codec-only helpers of the original are not part
// of the segmentation path and are not provided.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <type_traits>

namespace pcc {

template <typename T>
class Vec3 {
public:
  T data[3];

  Vec3() = default;
  Vec3(T v) : data{v, v, v} {}
  Vec3(T x, T y, T z) : data{x, y, z} {}
  template <typename U>
  Vec3(const Vec3<U>& o) : data{T(o[0]), T(o[1]), T(o[2])} {}

  T& operator[](size_t i) { return data[i]; }
  const T& operator[](size_t i) const { return data[i]; }
  T& x() { return data[0]; }
  T& y() { return data[1]; }
  T& z() { return data[2]; }
  const T& x() const { return data[0]; }
  const T& y() const { return data[1]; }
  const T& z() const { return data[2]; }
  size_t getElementCount() const { return 3; }

  T getNorm2() const { return data[0] * data[0] + data[1] * data[1] + data[2] * data[2]; }
  T getNorm1() const { return std::abs(data[0]) + std::abs(data[1]) + std::abs(data[2]); }
  double getNorm() const { return std::sqrt(double(getNorm2())); }
  void normalize()
  {
    const double n = getNorm();
    if (n != 0.0) *this /= n;
  }

  Vec3& operator=(const T v) { data[0] = data[1] = data[2] = v; return *this; }
  template <typename U>
  Vec3& operator+=(const Vec3<U>& o) { data[0] += o[0]; data[1] += o[1]; data[2] += o[2]; return *this; }
  template <typename U>
  Vec3& operator-=(const Vec3<U>& o) { data[0] -= o[0]; data[1] -= o[1]; data[2] -= o[2]; return *this; }
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
  Vec3& operator*=(const U a) { data[0] *= a; data[1] *= a; data[2] *= a; return *this; }
  template <typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
  Vec3& operator/=(const U a) { data[0] /= a; data[1] /= a; data[2] /= a; return *this; }
  Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }

  bool operator==(const Vec3& o) const { return data[0] == o[0] && data[1] == o[1] && data[2] == o[2]; }
  bool operator!=(const Vec3& o) const { return !(*this == o); }
  bool operator<(const Vec3& o) const
  {
    if (data[0] != o[0]) return data[0] < o[0];
    if (data[1] != o[1]) return data[1] < o[1];
    return data[2] < o[2];
  }
};

template <typename T, typename U>
inline Vec3<typename std::common_type<T, U>::type> operator+(const Vec3<T>& a, const Vec3<U>& b)
{
  return {a[0] + b[0], a[1] + b[1], a[2] + b[2]};
}
template <typename T, typename U>
inline Vec3<typename std::common_type<T, U>::type> operator-(const Vec3<T>& a, const Vec3<U>& b)
{
  return {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
}
// dot product, left to right
template <typename T, typename U>
inline typename std::common_type<T, U>::type operator*(const Vec3<T>& a, const Vec3<U>& b)
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
template <typename T, typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
inline Vec3<typename std::common_type<T, U>::type> operator*(const Vec3<T>& a, const U s)
{
  return {a[0] * s, a[1] * s, a[2] * s};
}
template <typename T, typename U, typename = typename std::enable_if<std::is_arithmetic<U>::value>::type>
inline Vec3<typename std::common_type<T, U>::type> operator/(const Vec3<T>& a, const U s)
{
  return {a[0] / s, a[1] / s, a[2] / s};
}

template <typename T>
struct Box3 {
  Vec3<T> min, max;
  Box3() = default;
  Box3(T lo, T hi) : min(lo), max(hi) {}
  bool contains(const Vec3<T>& p) const
  {
    return p[0] >= min[0] && p[0] <= max[0] && p[1] >= min[1] && p[1] <= max[1] && p[2] >= min[2] && p[2] <= max[2];
  }
  void insert(const Vec3<T>& p)
  {
    for (int k = 0; k < 3; ++k) {
      if (p[k] < min[k]) min[k] = p[k];
      if (p[k] > max[k]) max[k] = p[k];
    }
  }
};

typedef Vec3<int32_t> point_t;
typedef uint16_t attr_t;

}  // namespace pcc
