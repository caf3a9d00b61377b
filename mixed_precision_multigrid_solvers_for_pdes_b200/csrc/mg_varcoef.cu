// libmgb200: variable-coefficient operator  A u = -div(a grad u) + shift*u  (SURVEY 8f-1).
//
// The reference advertises -div(a grad u) = f (README.md:175) but ships no operator for it, so there is no reference
// arithmetic to mirror (parity unpinned); the discretisation is the standard conservative 5-point one with
// arithmetic-mean face coefficients of the nodal field a:
//     a_e = (a[i+1][j] + a[i][j])/2, a_w, a_n, a_s likewise,
//     (A u)[i][j] = (a_e (u_c - u_e) + a_w (u_c - u_w))/hx^2 + (a_n (u_c - u_n) + a_s (u_c - u_s))/hy^2 + shift*u_c.
// For a == 1 it reduces to -lap_h.  Strict (non-contracted) arithmetic so that oracle/np_oracle.py matches bit for bit.
#include "mg_common.cuh"

namespace mg {

constexpr int VBX = 128, VBY = 2;

template <typename T> struct VarScalars { T hx2, hy2, shift, omega, one_minus_omega; };

template <typename T>
__device__ __forceinline__ void faces(const T* a, int64_t lda, T& ae, T& aw, T& an, T& as) {
  using A = Strict<T>;
  ae = A::mul((T)0.5, A::add(a[lda], a[0]));   // first index = x: "east" is i+1
  aw = A::mul((T)0.5, A::add(a[-lda], a[0]));
  an = A::mul((T)0.5, A::add(a[1], a[0]));
  as = A::mul((T)0.5, A::add(a[-1], a[0]));
}

template <typename T>
__global__ void __launch_bounds__(VBX* VBY) varcoef_residual_kernel(const T* __restrict__ u, const T* __restrict__ f,
                                                                    const T* __restrict__ a, T* __restrict__ r, int nx,
                                                                    int ny, int64_t ldu, int64_t ldf, int64_t lda,
                                                                    int64_t ldr, VarScalars<T> s, int apply_only) {
  const int j = blockIdx.x * VBX + threadIdx.x, i = blockIdx.y * VBY + threadIdx.y;
  if (i >= nx || j >= ny) return;
  using A = Strict<T>;
  T au = (T)0;
  const bool interior = i > 0 && i < nx - 1 && j > 0 && j < ny - 1;
  if (interior) {
    const T* p = u + (int64_t)i * ldu + j;
    T ae, aw, an, as;
    faces<T>(a + (int64_t)i * lda + j, lda, ae, aw, an, as);
    const T x = A::div(A::add(A::mul(ae, A::sub(p[0], p[ldu])), A::mul(aw, A::sub(p[0], p[-ldu]))), s.hx2);
    const T y = A::div(A::add(A::mul(an, A::sub(p[0], p[1])), A::mul(as, A::sub(p[0], p[-1]))), s.hy2);
    au = A::add(A::add(x, y), A::mul(s.shift, p[0]));
  }
  r[(int64_t)i * ldr + j] = apply_only ? au : A::sub(f[(int64_t)i * ldf + j], au);  // r = f on the boundary
}

template <typename T>
__global__ void __launch_bounds__(VBX* VBY) varcoef_rbgs_kernel(T* __restrict__ u, const T* __restrict__ f,
                                                                const T* __restrict__ a, int nx, int ny, int64_t ldu,
                                                                int64_t ldf, int64_t lda, int colour, VarScalars<T> s) {
  const int jj = blockIdx.x * VBX + threadIdx.x, i = 1 + blockIdx.y * VBY + threadIdx.y;
  if (i >= nx - 1) return;
  const int j = 2 * jj + 1 + ((i + 1 + colour) & 1);
  if (j >= ny - 1) return;
  using A = Strict<T>;
  T* p = u + (int64_t)i * ldu + j;
  T ae, aw, an, as;
  faces<T>(a + (int64_t)i * lda + j, lda, ae, aw, an, as);
  const T nb = A::add(A::div(A::add(A::mul(ae, p[ldu]), A::mul(aw, p[-ldu])), s.hx2),
                      A::div(A::add(A::mul(an, p[1]), A::mul(as, p[-1])), s.hy2));
  const T diag = A::add(A::add(A::div(A::add(ae, aw), s.hx2), A::div(A::add(an, as), s.hy2)), s.shift);
  const T unew = A::div(A::add(f[(int64_t)i * ldf + j], nb), diag);
  p[0] = A::add(A::mul(s.one_minus_omega, p[0]), A::mul(s.omega, unew));
}

// Coarsest level: <= maxit x [one red-black GS sweep, residual, h-scaled L2 norm over all points], stop below tol --
// IterativeSolver.solve (solvers/base.py:258-285) with VariableCoefficientSmoother as the solver, in ONE launch of one
// block with the stopping test in the kernel (the host-driven loop costs a stream synchronisation per sweep).
// Same strict arithmetic as the two kernels above.
constexpr int VCS_THREADS = 256;
template <typename T>
__global__ void __launch_bounds__(VCS_THREADS) varcoef_coarse_solve_kernel(T* __restrict__ u, const T* __restrict__ f,
                                                                          const T* __restrict__ a, int nx, int ny,
                                                                          int64_t ldu, int64_t ldf, int64_t lda,
                                                                          VarScalars<T> s, double hxhy, double tol, int maxit,
                                                                          double* __restrict__ info) {
  using A = Strict<T>;
  __shared__ double red[32];
  const int n = nx * ny;
  int it = 1;
  double norm = 0.0;
  for (; it <= maxit; ++it) {
    for (int colour = 0; colour < 2; ++colour) {
      for (int k = threadIdx.x; k < n; k += VCS_THREADS) {
        const int i = k / ny, j = k - i * ny;
        if (i < 1 || i > nx - 2 || j < 1 || j > ny - 2 || ((i + j) & 1) != colour) continue;
        T* p = u + (int64_t)i * ldu + j;
        T ae, aw, an, as;
        faces<T>(a + (int64_t)i * lda + j, lda, ae, aw, an, as);
        const T nb = A::add(A::div(A::add(A::mul(ae, p[ldu]), A::mul(aw, p[-ldu])), s.hx2),
                            A::div(A::add(A::mul(an, p[1]), A::mul(as, p[-1])), s.hy2));
        const T diag = A::add(A::add(A::div(A::add(ae, aw), s.hx2), A::div(A::add(an, as), s.hy2)), s.shift);
        const T unew = A::div(A::add(f[(int64_t)i * ldf + j], nb), diag);
        p[0] = A::add(A::mul(s.one_minus_omega, p[0]), A::mul(s.omega, unew));
      }
      __syncthreads();
    }
    double acc = 0.0;
    for (int k = threadIdx.x; k < n; k += VCS_THREADS) {
      const int i = k / ny, j = k - i * ny;
      T v = f[(int64_t)i * ldf + j];
      if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
        const T* p = u + (int64_t)i * ldu + j;
        T ae, aw, an, as;
        faces<T>(a + (int64_t)i * lda + j, lda, ae, aw, an, as);
        const T x = A::div(A::add(A::mul(ae, A::sub(p[0], p[ldu])), A::mul(aw, A::sub(p[0], p[-ldu]))), s.hx2);
        const T y = A::div(A::add(A::mul(an, A::sub(p[0], p[1])), A::mul(as, A::sub(p[0], p[-1]))), s.hy2);
        v = A::sub(v, A::add(A::add(x, y), A::mul(s.shift, p[0])));
      }
      acc += (double)A::mul(v, v);
    }
    acc = block_reduce(acc, red);
    norm = sqrt(hxhy * acc);
    if (norm < tol) break;
  }
  if (info != nullptr && threadIdx.x == 0) {
    info[0] = (double)(it > maxit ? maxit : it);
    info[1] = norm;
  }
}

template <typename T>
static VarScalars<T> var_scalars(double hx, double hy, double shift, double omega) {
  VarScalars<T> s;
  s.hx2 = (T)pow(hx, 2.0);
  s.hy2 = (T)pow(hy, 2.0);
  s.shift = (T)shift;
  s.omega = (T)omega;
  s.one_minus_omega = (T)(1 - omega);
  return s;
}

}  // namespace mg

using namespace mg;

extern "C" {

int mg_varcoef_residual(const void* u, const void* f, const void* a, void* r, int nx, int ny, int64_t ld_u,
                        int64_t ld_f, int64_t ld_a, int64_t ld_r, double hx, double hy, double shift, int apply_only,
                        int dtype, void* stream) {
  if (!u || !a || !r || (!apply_only && !f) || nx < 3 || ny < 3 || ld_u < ny || ld_a < ny || ld_r < ny || hx <= 0 ||
      hy <= 0 || !(shift >= 0))
    return MG_ERR_BADARG;
  if (dtype != MG_F32 && dtype != MG_F64) return MG_ERR_DTYPE;
  const dim3 g((ny + VBX - 1) / VBX, (nx + VBY - 1) / VBY), b(VBX, VBY);
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    varcoef_residual_kernel<double><<<g, b, 0, st>>>((const double*)u, (const double*)f, (const double*)a, (double*)r, nx,
                                                     ny, ld_u, ld_f, ld_a, ld_r, var_scalars<double>(hx, hy, shift, 1.0),
                                                     apply_only);
  else
    varcoef_residual_kernel<float><<<g, b, 0, st>>>((const float*)u, (const float*)f, (const float*)a, (float*)r, nx, ny,
                                                    ld_u, ld_f, ld_a, ld_r, var_scalars<float>(hx, hy, shift, 1.0),
                                                    apply_only);
  return check_launch("mg_varcoef_residual");
}

int mg_varcoef_smooth_rbgs(void* u, const void* f, const void* a, int nx, int ny, int64_t ld_u, int64_t ld_f,
                           int64_t ld_a, double hx, double hy, double shift, double omega, int sweeps, int dtype,
                           void* stream) {
  if (!u || !f || !a || nx < 3 || ny < 3 || ld_u < ny || ld_f < ny || ld_a < ny || hx <= 0 || hy <= 0 || !(shift >= 0) ||
      sweeps < 0)
    return MG_ERR_BADARG;
  if (dtype != MG_F32 && dtype != MG_F64) return MG_ERR_DTYPE;
  const int nj = (ny - 2 + 1) / 2;
  const dim3 g((nj + VBX - 1) / VBX, (nx - 2 + VBY - 1) / VBY), b(VBX, VBY);
  cudaStream_t st = as_stream(stream);
  for (int k = 0; k < sweeps; ++k)
    for (int colour = 0; colour < 2; ++colour) {
      if (dtype == MG_F64)
        varcoef_rbgs_kernel<double><<<g, b, 0, st>>>((double*)u, (const double*)f, (const double*)a, nx, ny, ld_u, ld_f,
                                                     ld_a, colour, var_scalars<double>(hx, hy, shift, omega));
      else
        varcoef_rbgs_kernel<float><<<g, b, 0, st>>>((float*)u, (const float*)f, (const float*)a, nx, ny, ld_u, ld_f, ld_a,
                                                    colour, var_scalars<float>(hx, hy, shift, omega));
    }
  return check_launch("mg_varcoef_smooth_rbgs", 2 * sweeps);
}

int mg_varcoef_coarse_solve(void* u, const void* f, const void* a, int nx, int ny, int64_t ld_u, int64_t ld_f, int64_t ld_a,
                            double hx, double hy, double shift, double omega, double tolerance, int max_iterations,
                            double* info, int dtype, void* stream) {
  if (!u || !f || !a || nx < 3 || ny < 3 || ld_u < ny || ld_f < ny || ld_a < ny || hx <= 0 || hy <= 0 || !(shift >= 0) ||
      max_iterations < 1)
    return MG_ERR_BADARG;
  if (dtype != MG_F32 && dtype != MG_F64) return MG_ERR_DTYPE;
  if ((int64_t)nx * ny > 65 * 65) return MG_ERR_UNSUPPORTED;  // one block: coarsest grids only
  cudaStream_t st = as_stream(stream);
  if (dtype == MG_F64)
    varcoef_coarse_solve_kernel<double><<<1, VCS_THREADS, 0, st>>>((double*)u, (const double*)f, (const double*)a, nx, ny,
                                                                   ld_u, ld_f, ld_a, var_scalars<double>(hx, hy, shift, omega),
                                                                   hx * hy, tolerance, max_iterations, info);
  else
    varcoef_coarse_solve_kernel<float><<<1, VCS_THREADS, 0, st>>>((float*)u, (const float*)f, (const float*)a, nx, ny, ld_u,
                                                                  ld_f, ld_a, var_scalars<float>(hx, hy, shift, omega),
                                                                  hx * hy, tolerance, max_iterations, info);
  return check_launch("mg_varcoef_coarse_solve");
}

}  // extern "C"
