/*
 * mg_oracle.c  --  plain-C restatement of the reference multigrid hot path.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never linked into the product.
 *
 * Parity: PINNED through tests/test_oracle_golden.py (bit-for-bit against outputs of the
 * reference's own Python classes, tests/golden/.npz).
 *
 * Every function evaluates, per point, the same IEEE operations in the same order as the
 * reference loop it cites (compile with -ffp-contract=off: no FMA contraction).  OpenMP is
 * applied only where the reference's loop iterations are independent (one colour of the
 * red-black sweep, Jacobi, residual, transfers), so threading does not change a single bit.
 *
 * Arrays: row-major (nx, ny), boundary points included, dense (ld = ny).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(i, j) ((size_t)(i) * (size_t)ny + (size_t)(j))

#define DEFINE_ALL(T, SFX)                                                                                   \
  /* smoothers.py:187-193 (also :163-170, :72-82) */                                                         \
  static inline T relax_##SFX(T uc, T up, T dn, T rt, T lf, T rhs, T hx2, T hy2, T negdiag, T om1, T om) {   \
    T nb = (up + dn) / hx2 + (rt + lf) / hy2;                                                                \
    T unew = (rhs + nb) / negdiag;                                                                           \
    return om1 * uc + om * unew;                                                                             \
  }                                                                                                          \
  /* LaplacianOperator.apply, operators/laplacian.py:44-80 */                                                \
  void orc_apply_##SFX(const T* u, T* out, int nx, int ny, double hx, double hy, double coeff) {             \
    const T hx2 = (T)pow(hx, 2.0), hy2 = (T)pow(hy, 2.0), cc = (T)(2.0 / pow(hx, 2.0) + 2.0 / pow(hy, 2.0)); \
    const T c = (T)coeff;                                                                                    \
    _Pragma("omp parallel for schedule(static)") for (int i = 0; i < nx; ++i) {                              \
      for (int j = 0; j < ny; ++j) {                                                                         \
        T v = (T)0;                                                                                          \
        if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1)                                                      \
          v = c * ((u[IDX(i + 1, j)] + u[IDX(i - 1, j)]) / hx2 + (u[IDX(i, j + 1)] + u[IDX(i, j - 1)]) / hy2 - \
                   u[IDX(i, j)] * cc);                                                                       \
        out[IDX(i, j)] = v;                                                                                  \
      }                                                                                                      \
    }                                                                                                        \
  }                                                                                                          \
  /* LaplacianOperator.residual, operators/laplacian.py:105-124: r = f - A u (r = f on the boundary) */      \
  void orc_residual_##SFX(const T* u, const T* f, T* r, int nx, int ny, double hx, double hy, double coeff) { \
    const T hx2 = (T)pow(hx, 2.0), hy2 = (T)pow(hy, 2.0), cc = (T)(2.0 / pow(hx, 2.0) + 2.0 / pow(hy, 2.0)); \
    const T c = (T)coeff;                                                                                    \
    _Pragma("omp parallel for schedule(static)") for (int i = 0; i < nx; ++i) {                              \
      for (int j = 0; j < ny; ++j) {                                                                         \
        T v = (T)0;                                                                                          \
        if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1)                                                      \
          v = c * ((u[IDX(i + 1, j)] + u[IDX(i - 1, j)]) / hx2 + (u[IDX(i, j + 1)] + u[IDX(i, j - 1)]) / hy2 - \
                   u[IDX(i, j)] * cc);                                                                       \
        r[IDX(i, j)] = f[IDX(i, j)] - v;                                                                     \
      }                                                                                                      \
    }                                                                                                        \
  }                                                                                                          \
  /* GaussSeidelSmoother._red_black_sweep, solvers/smoothers.py:175-207 (in place) */                        \
  void orc_rbgs_##SFX(T* u, const T* f, int nx, int ny, double hx, double hy, double omega, int sweeps) {    \
    const double diag = -2.0 / pow(hx, 2.0) - 2.0 / pow(hy, 2.0);                                            \
    const T hx2 = (T)pow(hx, 2.0), hy2 = (T)pow(hy, 2.0), nd = (T)(-diag), om = (T)omega, om1 = (T)(1 - omega); \
    for (int s = 0; s < sweeps; ++s)                                                                         \
      for (int colour = 0; colour < 2; ++colour) {                                                           \
        _Pragma("omp parallel for schedule(static)") for (int i = 1; i < nx - 1; ++i) {                      \
          for (int j = 1 + ((i + 1 + colour) & 1); j < ny - 1; j += 2)                                       \
            u[IDX(i, j)] = relax_##SFX(u[IDX(i, j)], u[IDX(i + 1, j)], u[IDX(i - 1, j)], u[IDX(i, j + 1)],   \
                                       u[IDX(i, j - 1)], f[IDX(i, j)], hx2, hy2, nd, om1, om);               \
        }                                                                                                    \
      }                                                                                                      \
  }                                                                                                          \
  /* GaussSeidelSmoother._lexicographic_sweep, solvers/smoothers.py:153-173 (in place, sequential) */        \
  void orc_lexgs_##SFX(T* u, const T* f, int nx, int ny, double hx, double hy, double omega, int sweeps) {   \
    const double diag = -2.0 / pow(hx, 2.0) - 2.0 / pow(hy, 2.0);                                            \
    const T hx2 = (T)pow(hx, 2.0), hy2 = (T)pow(hy, 2.0), nd = (T)(-diag), om = (T)omega, om1 = (T)(1 - omega); \
    for (int s = 0; s < sweeps; ++s)                                                                         \
      for (int i = 1; i < nx - 1; ++i)                                                                       \
        for (int j = 1; j < ny - 1; ++j)                                                                     \
          u[IDX(i, j)] = relax_##SFX(u[IDX(i, j)], u[IDX(i + 1, j)], u[IDX(i - 1, j)], u[IDX(i, j + 1)],     \
                                     u[IDX(i, j - 1)], f[IDX(i, j)], hx2, hy2, nd, om1, om);                 \
  }                                                                                                          \
  /* JacobiSmoother.smooth, solvers/smoothers.py:41-86: result in u, tmp is scratch of the same size */      \
  void orc_jacobi_##SFX(T* u, T* tmp, const T* f, int nx, int ny, double hx, double hy, double omega,        \
                        int sweeps) {                                                                        \
    const double diag = -2.0 / pow(hx, 2.0) - 2.0 / pow(hy, 2.0);                                            \
    const T hx2 = (T)pow(hx, 2.0), hy2 = (T)pow(hy, 2.0), nd = (T)(-diag), om = (T)omega, om1 = (T)(1 - omega); \
    for (int s = 0; s < sweeps; ++s) {                                                                       \
      memcpy(tmp, u, sizeof(T) * (size_t)nx * ny);                                                           \
      _Pragma("omp parallel for schedule(static)") for (int i = 1; i < nx - 1; ++i) {                        \
        for (int j = 1; j < ny - 1; ++j)                                                                     \
          u[IDX(i, j)] = relax_##SFX(tmp[IDX(i, j)], tmp[IDX(i + 1, j)], tmp[IDX(i - 1, j)],                 \
                                     tmp[IDX(i, j + 1)], tmp[IDX(i, j - 1)], f[IDX(i, j)], hx2, hy2, nd, om1, om); \
      }                                                                                                      \
    }                                                                                                        \
  }                                                                                                          \
  /* RestrictionOperator._full_weighting_restriction / _injection / _half_weighting,                         \
     operators/transfer.py:83-148.  method: 0 fw, 1 injection, 2 hw.  Arithmetic in T. */                    \
  void orc_restrict_##SFX(const T* fine, T* coarse, int nxf, int nyf, int method) {                          \
    const int nxc = (nxf - 1) / 2 + 1, nyc = (nyf - 1) / 2 + 1, ny = nyf;                                    \
    _Pragma("omp parallel for schedule(static)") for (int i = 0; i < nxc; ++i) {                             \
      for (int j = 0; j < nyc; ++j) {                                                                        \
        const int fi = 2 * i, fj = 2 * j;                                                                    \
        T v = fine[IDX(fi, fj)];                                                                             \
        if (method != 1 && i > 0 && i < nxc - 1 && j > 0 && j < nyc - 1) {                                   \
          const T edges = fine[IDX(fi - 1, fj)] + fine[IDX(fi + 1, fj)] + fine[IDX(fi, fj - 1)] + fine[IDX(fi, fj + 1)]; \
          if (method == 0) {                                                                                 \
            const T corners = fine[IDX(fi - 1, fj - 1)] + fine[IDX(fi - 1, fj + 1)] + fine[IDX(fi + 1, fj - 1)] + \
                              fine[IDX(fi + 1, fj + 1)];                                                     \
            v = (T)(1.0 / 16.0) * corners + (T)(1.0 / 8.0) * edges + (T)(1.0 / 4.0) * fine[IDX(fi, fj)];     \
          } else {                                                                                           \
            v = (T)(1.0 / 8.0) * edges + (T)(1.0 / 2.0) * fine[IDX(fi, fj)];                                 \
          }                                                                                                  \
        }                                                                                                    \
        coarse[(size_t)i * nyc + j] = v;                                                                     \
      }                                                                                                      \
    }                                                                                                        \
  }                                                                                                          \
  /* ProlongationOperator._bilinear_prolongation / _injection, operators/transfer.py:217-267, incl. the      \
     guards that leave odd points of the last fine row / column at 0.  method: 0 bilinear, 1 injection. */   \
  void orc_prolong_##SFX(const T* c, T* fine, int nxc, int nyc, int method) {                                \
    const int nxf = 2 * (nxc - 1) + 1, nyf = 2 * (nyc - 1) + 1, ny = nyf;                                    \
    _Pragma("omp parallel for schedule(static)") for (int i = 0; i < nxf; ++i) {                             \
      for (int j = 0; j < nyf; ++j) {                                                                        \
        const int ic = i >> 1, jc = j >> 1;                                                                  \
        const T* p = c + (size_t)ic * nyc + jc;                                                              \
        T v = (T)0;                                                                                          \
        if (!(i & 1) && !(j & 1)) v = p[0];                                                                  \
        else if (method == 0) {                                                                              \
          if ((i & 1) && !(j & 1)) { if (j < nyf - 1) v = (T)0.5 * (p[0] + p[nyc]); }                        \
          else if (!(i & 1) && (j & 1)) { if (i < nxf - 1) v = (T)0.5 * (p[0] + p[1]); }                     \
          else v = (T)0.25 * (p[0] + p[1] + p[nyc] + p[nyc + 1]);                                            \
        }                                                                                                    \
        fine[IDX(i, j)] = v;                                                                                 \
      }                                                                                                      \
    }                                                                                                        \
  }                                                                                                          \
  /* sum over all points of x*x (square in T, accumulate in double, fixed row-block order):                  \
     the argument of Grid.l2_norm, core/grid.py:174-187 */                                                   \
  double orc_sumsq_##SFX(const T* x, int nx, int ny) {                                                       \
    double total = 0.0;                                                                                      \
    _Pragma("omp parallel for schedule(static) reduction(+ : total)") for (int i = 0; i < nx; ++i) {         \
      double acc = 0.0;                                                                                      \
      for (int j = 0; j < ny; ++j) { const T v = x[IDX(i, j)]; acc += (double)(v * v); }                     \
      total += acc;                                                                                          \
    }                                                                                                        \
    return total;                                                                                            \
  }                                                                                                          \
  void orc_axpy_##SFX(T* y, const T* x, size_t n) {                                                          \
    _Pragma("omp parallel for schedule(static)") for (size_t k = 0; k < n; ++k) y[k] = y[k] + x[k];          \
  }

DEFINE_ALL(double, f64)
DEFINE_ALL(float, f32)

int orc_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
