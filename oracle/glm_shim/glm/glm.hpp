// Minimal stand-in for the parts of GLM (un-vendored, unpinned dependency of ohm: CMakeLists.txt:106, vcpkg.json)
// that ohm's CPU ray-integration path touches.  TEST INFRASTRUCTURE ONLY: it exists so that the reference's own
// sources can be compiled, unmodified and from where they lie under /root/reference, into oracle/_ref (see
// oracle/Makefile).  Semantics follow GLM 0.9.9's generic (non-SIMD) code paths:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z        (detail::compute_dot<vec<3,...>>)
//   length(v)     = sqrt(dot(v,v)),  length2(v) = dot(v,v)
//   normalize(v)  = v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1 / sqrt(x)
//   vec op= scalar converts the scalar to the vector's value type first.
#ifndef OHM_ORACLE_GLM_SHIM_HPP
#define OHM_ORACLE_GLM_SHIM_HPP

#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <utility>

namespace glm
{
enum qualifier
{
  packed_highp,
  highp = packed_highp,
  defaultp = highp
};
typedef qualifier precision;
typedef int length_t;

template <length_t L, typename T, qualifier Q = defaultp>
struct vec;

template <typename T, qualifier Q>
struct vec<2, T, Q>
{
  typedef T value_type;
  T x, y;
  constexpr vec() : x(0), y(0) {}
  constexpr explicit vec(T s) : x(s), y(s) {}
  constexpr vec(T x_, T y_) : x(x_), y(y_) {}
  template <typename U, qualifier P>
  constexpr vec(const vec<2, U, P> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
  T &operator[](length_t i) { return (&x)[i]; }
  constexpr const T &operator[](length_t i) const { return (&x)[i]; }
  static constexpr length_t length() { return 2; }
};

template <typename T, qualifier Q>
struct vec<3, T, Q>
{
  typedef T value_type;
  T x, y, z;
  constexpr vec() : x(0), y(0), z(0) {}
  constexpr explicit vec(T s) : x(s), y(s), z(s) {}
  template <typename A, typename B, typename C>
  constexpr vec(A x_, B y_, C z_) : x(static_cast<T>(x_)), y(static_cast<T>(y_)), z(static_cast<T>(z_)) {}
  template <typename U, qualifier P>
  constexpr vec(const vec<3, U, P> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
  template <typename U, qualifier P>
  constexpr explicit vec(const vec<4, U, P> &v);
  T &operator[](length_t i) { return (&x)[i]; }
  constexpr const T &operator[](length_t i) const { return (&x)[i]; }
  static constexpr length_t length() { return 3; }

  template <typename U>
  vec &operator+=(const vec<3, U, Q> &v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); z += static_cast<T>(v.z); return *this; }
  template <typename U>
  vec &operator-=(const vec<3, U, Q> &v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); z -= static_cast<T>(v.z); return *this; }
  template <typename U>
  vec &operator*=(const vec<3, U, Q> &v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); z *= static_cast<T>(v.z); return *this; }
  template <typename U>
  vec &operator/=(const vec<3, U, Q> &v) { x /= static_cast<T>(v.x); y /= static_cast<T>(v.y); z /= static_cast<T>(v.z); return *this; }
  template <typename U>
  vec &operator+=(U s) { x += static_cast<T>(s); y += static_cast<T>(s); z += static_cast<T>(s); return *this; }
  template <typename U>
  vec &operator-=(U s) { x -= static_cast<T>(s); y -= static_cast<T>(s); z -= static_cast<T>(s); return *this; }
  template <typename U>
  vec &operator*=(U s) { x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s); return *this; }
  template <typename U>
  vec &operator/=(U s) { x /= static_cast<T>(s); y /= static_cast<T>(s); z /= static_cast<T>(s); return *this; }
};

template <typename T, qualifier Q>
struct vec<4, T, Q>
{
  typedef T value_type;
  T x, y, z, w;
  constexpr vec() : x(0), y(0), z(0), w(0) {}
  constexpr explicit vec(T s) : x(s), y(s), z(s), w(s) {}
  template <typename A, typename B, typename C, typename D>
  constexpr vec(A x_, B y_, C z_, D w_) : x(static_cast<T>(x_)), y(static_cast<T>(y_)), z(static_cast<T>(z_)), w(static_cast<T>(w_)) {}
  template <typename U, qualifier P, typename D>
  constexpr vec(const vec<3, U, P> &v, D w_) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(w_)) {}
  template <typename U, qualifier P>
  constexpr vec(const vec<4, U, P> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(v.w)) {}
  T &operator[](length_t i) { return (&x)[i]; }
  constexpr const T &operator[](length_t i) const { return (&x)[i]; }
  static constexpr length_t length() { return 4; }
};

template <typename T, qualifier Q>
template <typename U, qualifier P>
constexpr vec<3, T, Q>::vec(const vec<4, U, P> &v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

template <typename T, qualifier Q = defaultp> using tvec2 = vec<2, T, Q>;
template <typename T, qualifier Q = defaultp> using tvec3 = vec<3, T, Q>;
template <typename T, qualifier Q = defaultp> using tvec4 = vec<4, T, Q>;

typedef vec<2, double> dvec2;
typedef vec<3, double> dvec3;
typedef vec<4, double> dvec4;
typedef vec<2, float> vec2;
typedef vec<3, float> vec3;
typedef vec<4, float> vec4;
typedef vec<2, int> ivec2;
typedef vec<3, int> ivec3;
typedef vec<4, int> ivec4;
typedef vec<3, int32_t> i32vec3;
typedef vec<3, uint32_t> u32vec3;
typedef vec<3, unsigned> uvec3;
typedef vec<3, int16_t> i16vec3;
typedef vec<4, int16_t> i16vec4;
typedef vec<3, uint8_t> u8vec3;
typedef vec<4, uint8_t> u8vec4;
typedef vec<3, int8_t> i8vec3;
typedef vec<3, bool> bvec3;
typedef vec<4, bool> bvec4;
typedef vec<2, bool> bvec2;

// -- component-wise arithmetic ------------------------------------------------------------------------------
#define OHM_GLM_SHIM_BINOP(op)                                                                                   \
  template <typename T, qualifier Q>                                                                             \
  constexpr vec<3, T, Q> operator op(const vec<3, T, Q> &a, const vec<3, T, Q> &b)                               \
  {                                                                                                              \
    return vec<3, T, Q>(a.x op b.x, a.y op b.y, a.z op b.z);                                                     \
  }                                                                                                              \
  template <typename T, qualifier Q>                                                                             \
  constexpr vec<3, T, Q> operator op(const vec<3, T, Q> &a, T s)                                                 \
  {                                                                                                              \
    return vec<3, T, Q>(a.x op s, a.y op s, a.z op s);                                                           \
  }                                                                                                              \
  template <typename T, qualifier Q>                                                                             \
  constexpr vec<3, T, Q> operator op(T s, const vec<3, T, Q> &b)                                                 \
  {                                                                                                              \
    return vec<3, T, Q>(s op b.x, s op b.y, s op b.z);                                                           \
  }                                                                                                              \
  template <typename T, qualifier Q>                                                                             \
  constexpr vec<2, T, Q> operator op(const vec<2, T, Q> &a, const vec<2, T, Q> &b)                               \
  {                                                                                                              \
    return vec<2, T, Q>(a.x op b.x, a.y op b.y);                                                                 \
  }                                                                                                              \
  template <typename T, qualifier Q>                                                                             \
  constexpr vec<4, T, Q> operator op(const vec<4, T, Q> &a, const vec<4, T, Q> &b)                               \
  {                                                                                                              \
    return vec<4, T, Q>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w);                                         \
  }
OHM_GLM_SHIM_BINOP(+)
OHM_GLM_SHIM_BINOP(-)
OHM_GLM_SHIM_BINOP(*)
OHM_GLM_SHIM_BINOP(/)
#undef OHM_GLM_SHIM_BINOP

template <typename T, qualifier Q>
constexpr vec<3, T, Q> operator-(const vec<3, T, Q> &a)
{
  return vec<3, T, Q>(-a.x, -a.y, -a.z);
}
template <length_t L, typename T, qualifier Q>
constexpr bool operator==(const vec<L, T, Q> &a, const vec<L, T, Q> &b)
{
  for (length_t i = 0; i < L; ++i)
  {
    if (!(a[i] == b[i]))
    {
      return false;
    }
  }
  return true;
}
template <length_t L, typename T, qualifier Q>
constexpr bool operator!=(const vec<L, T, Q> &a, const vec<L, T, Q> &b)
{
  return !(a == b);
}

// -- functions ---------------------------------------------------------------------------------------------
template <typename T, qualifier Q>
constexpr T dot(const vec<3, T, Q> &a, const vec<3, T, Q> &b)
{
  return (a.x * b.x + a.y * b.y) + a.z * b.z;
}
template <typename T, qualifier Q>
constexpr T dot(const vec<2, T, Q> &a, const vec<2, T, Q> &b)
{
  return a.x * b.x + a.y * b.y;
}
template <typename T, qualifier Q>
inline T length2(const vec<3, T, Q> &v)
{
  return dot(v, v);
}
template <typename T, qualifier Q>
inline T length(const vec<3, T, Q> &v)
{
  return std::sqrt(dot(v, v));
}
template <typename T, qualifier Q>
inline T distance(const vec<3, T, Q> &a, const vec<3, T, Q> &b)
{
  return length(b - a);
}
template <typename T>
inline T inversesqrt(T x)
{
  return static_cast<T>(1) / std::sqrt(x);
}
template <typename T, qualifier Q>
inline vec<3, T, Q> normalize(const vec<3, T, Q> &v)
{
  return v * inversesqrt(dot(v, v));
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> cross(const vec<3, T, Q> &x, const vec<3, T, Q> &y)
{
  return vec<3, T, Q>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
template <typename T, qualifier Q>
inline vec<3, T, Q> abs(const vec<3, T, Q> &v)
{
  return vec<3, T, Q>(std::abs(v.x), std::abs(v.y), std::abs(v.z));
}
template <typename T, qualifier Q>
inline vec<3, T, Q> sqrt(const vec<3, T, Q> &v)
{
  return vec<3, T, Q>(std::sqrt(v.x), std::sqrt(v.y), std::sqrt(v.z));
}
template <typename T, qualifier Q>
inline vec<3, T, Q> floor(const vec<3, T, Q> &v)
{
  return vec<3, T, Q>(std::floor(v.x), std::floor(v.y), std::floor(v.z));
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> max(const vec<3, T, Q> &a, const vec<3, T, Q> &b)
{
  return vec<3, T, Q>(a.x < b.x ? b.x : a.x, a.y < b.y ? b.y : a.y, a.z < b.z ? b.z : a.z);
}
template <typename T, qualifier Q>
constexpr vec<3, T, Q> min(const vec<3, T, Q> &a, const vec<3, T, Q> &b)
{
  return vec<3, T, Q>(b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z);
}
template <typename T>
constexpr T max(T a, T b)
{
  return a < b ? b : a;
}
template <typename T>
constexpr T min(T a, T b)
{
  return b < a ? b : a;
}
template <typename T, qualifier Q>
inline vec<3, bool, Q> isnan(const vec<3, T, Q> &v)
{
  return vec<3, bool, Q>(std::isnan(v.x), std::isnan(v.y), std::isnan(v.z));
}
template <typename T, qualifier Q>
inline vec<3, bool, Q> isinf(const vec<3, T, Q> &v)
{
  return vec<3, bool, Q>(std::isinf(v.x), std::isinf(v.y), std::isinf(v.z));
}
#define OHM_GLM_SHIM_CMP(name, op)                                                                  \
  template <typename T, qualifier Q>                                                                \
  constexpr vec<3, bool, Q> name(const vec<3, T, Q> &a, const vec<3, T, Q> &b)                      \
  {                                                                                                 \
    return vec<3, bool, Q>(a.x op b.x, a.y op b.y, a.z op b.z);                                     \
  }
OHM_GLM_SHIM_CMP(equal, ==)
OHM_GLM_SHIM_CMP(notEqual, !=)
OHM_GLM_SHIM_CMP(lessThan, <)
OHM_GLM_SHIM_CMP(lessThanEqual, <=)
OHM_GLM_SHIM_CMP(greaterThan, >)
OHM_GLM_SHIM_CMP(greaterThanEqual, >=)
#undef OHM_GLM_SHIM_CMP
template <length_t L, qualifier Q>
constexpr bool any(const vec<L, bool, Q> &v)
{
  for (length_t i = 0; i < L; ++i)
  {
    if (v[i])
    {
      return true;
    }
  }
  return false;
}
template <length_t L, qualifier Q>
constexpr bool all(const vec<L, bool, Q> &v)
{
  for (length_t i = 0; i < L; ++i)
  {
    if (!v[i])
    {
      return false;
    }
  }
  return true;
}

template <length_t L, typename T, qualifier Q>
inline const T *value_ptr(const vec<L, T, Q> &v)
{
  return &v.x;
}
template <length_t L, typename T, qualifier Q>
inline T *value_ptr(vec<L, T, Q> &v)
{
  return &v.x;
}

// -- matrix / quaternion types: declarations only (they appear in signatures of headers on the path, e.g.
//    ohm/CovarianceVoxel.h; no translation unit compiled into oracle/_ref evaluates them) -------------------------
template <length_t C, length_t R, typename T, qualifier Q = defaultp>
struct mat
{
  vec<R, T, Q> value[C];
  constexpr mat() : value{} {}
  constexpr explicit mat(T s) : value{}
  {
    for (length_t i = 0; i < C && i < R; ++i)
    {
      value[i][i] = s;
    }
  }
  vec<R, T, Q> &operator[](length_t i) { return value[i]; }
  constexpr const vec<R, T, Q> &operator[](length_t i) const { return value[i]; }
};
// column-major, m[column][row], as GLM
template <length_t N, typename T, qualifier Q>
inline mat<N, N, T, Q> transpose(const mat<N, N, T, Q> &m)
{
  mat<N, N, T, Q> r;
  for (length_t c = 0; c < N; ++c)
  {
    for (length_t k = 0; k < N; ++k)
    {
      r[c][k] = m[k][c];
    }
  }
  return r;
}
template <length_t N, typename T, qualifier Q>
inline mat<N, N, T, Q> operator*(const mat<N, N, T, Q> &a, const mat<N, N, T, Q> &b)
{
  mat<N, N, T, Q> r;
  for (length_t c = 0; c < N; ++c)
  {
    for (length_t row = 0; row < N; ++row)
    {
      T sum = 0;
      for (length_t k = 0; k < N; ++k)
      {
        sum += a[k][row] * b[c][k];
      }
      r[c][row] = sum;
    }
  }
  return r;
}
typedef mat<3, 3, double> dmat3;
typedef mat<4, 4, double> dmat4;
typedef mat<3, 3, float> mat3;
typedef mat<4, 4, float> mat4;
template <typename T, qualifier Q = defaultp>
struct qua
{
  T x, y, z, w;
  constexpr qua() : x(0), y(0), z(0), w(1) {}
  constexpr qua(T w_, T x_, T y_, T z_) : x(x_), y(y_), z(z_), w(w_) {}
};
typedef qua<double> dquat;
typedef qua<float> quat;

template <typename T>
constexpr T pi()
{
  return static_cast<T>(3.14159265358979323846264338327950288);
}
}  // namespace glm

#endif  // OHM_ORACLE_GLM_SHIM_HPP
