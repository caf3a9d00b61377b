// Shared device/host helpers for libmgb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/mgb200.h"

namespace mg {

// ------------------------------------------------------------------------------------------
// Strict IEEE arithmetic: round-to-nearest intrinsics are never contracted into FMAs, so a
// kernel written with these reproduces NumPy's elementwise results bit for bit.
// ------------------------------------------------------------------------------------------
template <typename T> struct Strict;
template <> struct Strict<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Strict<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// Scalars of the 5-point operator, prepared on the host exactly as the reference's Python
// expressions evaluate them (double arithmetic, then one rounding to T).
template <typename T> struct StencilScalars {
  T hx2;       // hx**2                      (laplacian.py:74, smoothers.py:188)
  T hy2;       // hy**2
  T neg_diag;  // -(-2/hx**2 - 2/hy**2)      (smoothers.py:141,192)
  T cc;        // 2/hx**2 + 2/hy**2          (laplacian.py:76)
  T omega;     // relaxation parameter
  T one_minus_omega;  // (1 - omega)         (smoothers.py:193)
  T coeff;     // LaplacianOperator.coefficient
  T shift;     // Helmholtz shift lambda >= 0: operator coeff*lap_h + lambda (0 for the reference's Poisson path)
  // reciprocal forms for the fast (vector/fused) kernels
  T ihx2, ihy2, inv_neg_diag;
  // 1 when hx2, hy2 and neg_diag are exact powers of two: multiplying by the reciprocal then equals the
  // reference's division bit for bit, and the strict kernels may skip the (slow, high-latency) fp64 divisions
  int recip_exact;
};

inline bool is_pow2(double x) {
  int e;
  return x > 0 && frexp(x, &e) == 0.5;
}

template <typename T>
inline StencilScalars<T> make_scalars(double hx, double hy, double omega, double coeff, double shift = 0.0) {
  StencilScalars<T> s;
  const double hx2 = pow(hx, 2.0), hy2 = pow(hy, 2.0);  // same libm pow CPython's float ** uses
  const double diag = -2.0 / hx2 - 2.0 / hy2;
  s.hx2 = (T)hx2;
  s.hy2 = (T)hy2;
  s.neg_diag = (T)(-diag + shift);  // diagonal of -lap_h + lambda
  s.shift = (T)shift;
  s.cc = (T)(2.0 / hx2 + 2.0 / hy2);
  s.omega = (T)omega;
  s.one_minus_omega = (T)(1 - omega);
  s.coeff = (T)coeff;
  s.ihx2 = (T)1 / s.hx2;
  s.ihy2 = (T)1 / s.hy2;
  s.inv_neg_diag = (T)1 / s.neg_diag;
  s.recip_exact = (is_pow2((double)s.hx2) && is_pow2((double)s.hy2) && is_pow2((double)s.neg_diag)) ? 1 : 0;
  return s;
}

// One point of the smoother, exactly smoothers.py:187-193 (also :163-170, :72-82).
template <typename T>
__device__ __forceinline__ T relax_strict(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T rhs) {
  using A = Strict<T>;
  T unew;
  if (s.recip_exact) {  // same bits as the divisions below (power-of-two scalings are exact)
    const T nb = A::add(A::mul(A::add(up, dn), s.ihx2), A::mul(A::add(rt, lf), s.ihy2));
    unew = A::mul(A::add(rhs, nb), s.inv_neg_diag);
  } else {
    const T nb = A::add(A::div(A::add(up, dn), s.hx2), A::div(A::add(rt, lf), s.hy2));
    unew = A::div(A::add(rhs, nb), s.neg_diag);
  }
  return A::add(A::mul(s.one_minus_omega, uc), A::mul(s.omega, unew));
}

// coefficient * 5-point Laplacian at one interior point, exactly laplacian.py:73-77.
template <typename T>
__device__ __forceinline__ T apply_strict(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf) {
  using A = Strict<T>;
  const T t = s.recip_exact
                  ? A::sub(A::add(A::mul(A::add(up, dn), s.ihx2), A::mul(A::add(rt, lf), s.ihy2)), A::mul(uc, s.cc))
                  : A::sub(A::add(A::div(A::add(up, dn), s.hx2), A::div(A::add(rt, lf), s.hy2)), A::mul(uc, s.cc));
  const T au = A::mul(s.coeff, t);
  return s.shift != (T)0 ? A::add(au, A::mul(s.shift, uc)) : au;
}

// ------------------------------------------------------------------------------------------
// Deterministic block reduction (fixed shuffle tree + fixed smem order).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// All threads of the block must call; result valid in every thread. `red` = 32 doubles of smem.
template <bool MAX = false>
__device__ __forceinline__ double block_reduce(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  double x = (lane < nw) ? red[lane] : (MAX ? -1.0 : 0.0);
  x = MAX ? warp_max(x) : warp_sum(x);
  return x;
}

// argument check of the C-ABI entry points
#define MG_REQUIRE(cond)               \
  do {                                 \
    if (!(cond)) return MG_ERR_BADARG; \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
int check_launch(const char* what, int launches = 1);  // defined in mg_basic.cu; also counts kernel launches
int sm_count();
// out[0] = sum of partials[0..n) in a fixed order (deterministic); defined in mg_basic.cu
void reduce_partials_sum(const double* partials, int n, double* out, cudaStream_t st);

}  // namespace mg
