/* f2c_rt.h -- run-time support for the C++ that oracle/f2cpp.py generates from the
 * reference's Fortran (TEST INFRASTRUCTURE ONLY; see oracle/f2cpp.py).
 *
 *   FArr<T,R>   Fortran array: column-major, per-dimension lower bound, ALLOCATE semantics.
 *               Fresh storage is zero-filled (Fortran leaves it undefined; zero is what a
 *               fresh page gives and it makes the run deterministic).  -DF2C_BOUNDS checks
 *               every subscript and aborts with the array name, like `gfortran -fcheck=bounds`.
 *   f2c_max/min MAX / MIN as gfortran expands them for non-NaN operands.
 *   f2c_powi    real ** integer: the square-and-multiply sequence of libgcc's __powisf2,
 *               which is what gfortran calls (x**2 == x*x, x**4 == (x*x)*(x*x)).
 *   f2c_ipow    integer ** integer.
 */
#pragma once
#include <cmath>
#ifdef F2C_PORTABLE_MATH
/* Build variant libh9ref_pk.so: EXP / LOG / real**real go to the portable, bit-reproducible
 * double-precision kernels that the GPU's exact mode uses (hybrid9_b200/csrc/h9_physics.h,
 * h9::MathExact) instead of glibc's.  With the same three functions on both sides, the GPU's
 * exact mode must reproduce the translated reference BIT FOR BIT (tests/test_gpu_vs_ref_bitwise.py);
 * only these functions are taken from that header, no physics. */
#include "../hybrid9_b200/csrc/h9_physics.h"
#define powf(a, b) h9::MathExact::pow((a), (b))
#define expf(a) h9::MathExact::exp((a))
#define logf(a) h9::MathExact::log((a))
#endif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>

template <class T, int R>
struct FArr {
  const char* name;
  T* p = nullptr;
  long lo[R], n[R], stride[R];
  long total = 0;
  explicit FArr(const char* nm) : name(nm) {
    for (int d = 0; d < R; ++d) lo[d] = 1, n[d] = 0, stride[d] = 0;
  }
  FArr(const FArr&) = delete;
  FArr& operator=(const FArr&) = delete;
  ~FArr() { std::free(p); }
  bool allocated() const { return p != nullptr; }
  void alloc(std::initializer_list<long> los, std::initializer_list<long> his) {
    if (p) { /* ALLOCATE of an allocated array is an error in Fortran */
      std::fprintf(stderr, "f2c: ALLOCATE of already allocated array %s\n", name);
      std::abort();
    }
    auto l = los.begin();
    auto h = his.begin();
    long s = 1;
    for (int d = 0; d < R; ++d, ++l, ++h) {
      lo[d] = *l;
      n[d] = (*h >= *l) ? (*h - *l + 1) : 0;
      stride[d] = s;
      s *= n[d];
    }
    total = s;
    p = static_cast<T*>(std::calloc(s > 0 ? (size_t)s : 1, sizeof(T)));
    if (!p) {
      std::fprintf(stderr, "f2c: out of memory allocating %s\n", name);
      std::abort();
    }
  }
  void dealloc() {
    std::free(p);
    p = nullptr;
    total = 0;
  }
  long lb(int d) const { return lo[d]; }
  long ub(int d) const { return lo[d] + n[d] - 1; }
  long size() const { return total; }
  T* data() { return p; }
  const T* data() const { return p; }
#ifdef F2C_BOUNDS
  void check(int d, long i) const {
    if (!p || i < lo[d] || i >= lo[d] + n[d]) {
      std::fprintf(stderr, "f2c: subscript %d of %s out of bounds: %ld not in [%ld, %ld]%s\n",
                   d + 1, name, i, lo[d], lo[d] + n[d] - 1, p ? "" : " (not allocated)");
      std::abort();
    }
  }
#else
  void check(int, long) const {}
#endif
  template <class... I>
  T& operator()(I... idx) {
    static_assert(sizeof...(I) == R, "rank mismatch");
    const long ii[R] = {(long)idx...};
    check(0, ii[0]);
    long off = ii[0] - lo[0]; /* the first dimension is contiguous (column-major) */
    for (int d = 1; d < R; ++d) {
      check(d, ii[d]);
      off += (ii[d] - lo[d]) * stride[d];
    }
    return p[off];
  }
};

static inline float f2c_max(float a, float b) { return (b > a) ? b : a; }
static inline float f2c_min(float a, float b) { return (b < a) ? b : a; }
static inline int f2c_max(int a, int b) { return (b > a) ? b : a; }
static inline int f2c_min(int a, int b) { return (b < a) ? b : a; }

static inline float f2c_powi(float x, int m) {
  unsigned int n = (m < 0) ? -(unsigned int)m : (unsigned int)m;
  float y = (n % 2) ? x : 1.0f;
  while (n >>= 1) {
    x = x * x;
    if (n % 2) y = y * x;
  }
  return (m < 0) ? 1.0f / y : y;
}

static inline int f2c_ipow(int x, int m) {
  if (m < 0) return (x == 1) ? 1 : (x == -1 ? ((m % 2) ? -1 : 1) : 0);
  int y = 1;
  while (m-- > 0) y *= x;
  return y;
}
