// libmgb200: the coarse end of a cycle in ONE launch (one CTA, everything in shared memory).
//
// Below ~129^2 a level is a few microseconds of work but every kernel launch, TMA pipeline fill and
// grid-wide drain costs about as much, and a W-cycle visits level l 2^l times (8192 coarsest-grid solves
// per W-cycle at 32769^2).  This kernel runs the complete sub-cycle (V, W or F recursion of
// solvers/multigrid.py:253-337) for the levels that fit in 227 KB of shared memory:
//   load u, f of the entry level -> [pre-smooth, residual + full weighting, recurse, prolong + correct,
//   post-smooth] per level, lexicographic-GS solve to tolerance on the coarsest (solvers/base.py:258-285)
//   -> store u.
// Arithmetic per point is the same as in the streaming kernel (relax_fast / residual_fast, reference operation
// order for the transfers), so results do not depend on which kernel handles a level when the spacings are
// powers of two; the coarsest solve uses the strict (division) arithmetic of mg_coarse_solve_lexgs.
// The coarsest level may be held in a wider type TC than the others (fp32 levels, fp64 coarsest: the
// reference never converts the coarsest level, multigrid.py:270-272).
#include <string.h>
#include "mg_common.cuh"
#include "mg_stream.cuh"

namespace mg {
namespace small {

constexpr int MAXLEV = 8;
constexpr int THREADS = 1024;

template <typename T, typename TC> struct Params {
  int nlev;
  int nx[MAXLEV], ny[MAXLEV];
  unsigned off_u[MAXLEV], off_f[MAXLEV];  // byte offsets into dynamic shared memory
  StencilScalars<T> sc[MAXLEV];           // levels 0 .. nlev-2
  StencilScalars<TC> scc;                 // coarsest level (strict arithmetic)
  double hxhy_c;                          // hx*hy of the coarsest level (norm scaling)
  double ctol;
  int cmaxit;
  int cycle;  // 0 V, 1 W, 2 F
  int pre, post;
  int u_zero;
  int iso1;  // hx == hy and omega == 1: the streaming kernel's 5-instruction point update (same bits as there)
  T* u;
  const T* f;
  int64_t ld_u, ld_f;
  double* info;  // optional: {coarse sweeps of the last coarse solve, its norm}
};

using stream::relax_fast;
using stream::relax_iso1;
using stream::residual_fast;
using stream::residual_iso;

template <typename T>
__device__ __forceinline__ T resid_at(const T* u, const T* f, int nx, int ny, int i, int j, const StencilScalars<T>& s,
                                      bool iso1) {
  const T fv = f[i * ny + j];
  if (i == 0 || i == nx - 1 || j == 0 || j == ny - 1) return fv;  // r = f on the boundary (laplacian.py:64,117)
  const T* p = u + i * ny + j;
  return iso1 ? residual_iso<T>(s, p[0], p[ny], p[-ny], p[1], p[-1], fv)
              : residual_fast<T>(s, p[0], p[ny], p[-ny], p[1], p[-1], fv);
}

template <typename T>
__device__ void smooth(T* u, const T* f, int nx, int ny, const StencilScalars<T>& s, int sweeps, bool iso1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = 0; k < sweeps; ++k)
    for (int c = 0; c < 2; ++c) {
      for (int i = 1 + warp; i <= nx - 2; i += THREADS / 32)      // rows over warps, columns over lanes
        for (int j = 1 + ((i + 1 + c) & 1) + 2 * lane; j <= ny - 2; j += 64) {
          T* p = u + i * ny + j;
          p[0] = iso1 ? relax_iso1<T>(s, p[ny], p[-ny], p[1], p[-1], f[i * ny + j])
                      : relax_fast<T>(s, p[0], p[ny], p[-ny], p[1], p[-1], f[i * ny + j]);
        }
      __syncthreads();
    }
}

// f_c = R(f - A u): injection on the coarse boundary, full weighting inside, reference summation order
template <typename T, typename TO>
__device__ void restrict_residual(const T* u, const T* f, int nx, int ny, const StencilScalars<T>& s, TO* fc, int nxc,
                                  int nyc, bool iso1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int I = warp; I < nxc; I += THREADS / 32)
   for (int J = lane; J < nyc; J += 32) {
    const int idx = I * nyc + J, i = 2 * I, j = 2 * J;
    T v;
    if (I == 0 || I == nxc - 1 || J == 0 || J == nyc - 1) {
      v = resid_at<T>(u, f, nx, ny, i, j, s, iso1);
    } else {
      const T nw = resid_at<T>(u, f, nx, ny, i - 1, j - 1, s, iso1), ne = resid_at<T>(u, f, nx, ny, i - 1, j + 1, s, iso1);
      const T sw = resid_at<T>(u, f, nx, ny, i + 1, j - 1, s, iso1), se = resid_at<T>(u, f, nx, ny, i + 1, j + 1, s, iso1);
      const T n_ = resid_at<T>(u, f, nx, ny, i - 1, j, s, iso1), s_ = resid_at<T>(u, f, nx, ny, i + 1, j, s, iso1);
      const T w_ = resid_at<T>(u, f, nx, ny, i, j - 1, s, iso1), e_ = resid_at<T>(u, f, nx, ny, i, j + 1, s, iso1);
      const T corners = ((nw + ne) + sw) + se;
      const T edges = ((n_ + s_) + w_) + e_;
      v = ((T)0.0625 * corners + (T)0.125 * edges) + (T)0.25 * resid_at<T>(u, f, nx, ny, i, j, s, iso1);
    }
    fc[idx] = (TO)v;
  }
  __syncthreads();
}

// u += P e_c, bilinear with the reference's last-row / last-column treatment (transfer.py:234-267)
template <typename T, typename TI>
__device__ void prolong_add(T* u, int nx, int ny, const TI* ec, int nyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < nx; i += THREADS / 32)
   for (int j = lane; j < ny; j += 32) {
    const int idx = i * ny + j;
    const TI* c = ec + (i >> 1) * nyc + (j >> 1);
    const bool oi = i & 1, oj = j & 1;
    T v = (T)0;
    if (!oi && !oj) v = (T)c[0];
    else if (oi && !oj) { if (j < ny - 1) v = (T)0.5 * ((T)c[0] + (T)c[nyc]); }
    else if (!oi && oj) { if (i < nx - 1) v = (T)0.5 * ((T)c[0] + (T)c[1]); }
    else v = (T)0.25 * ((((T)c[0] + (T)c[1]) + (T)c[nyc]) + (T)c[nyc + 1]);
    u[idx] += v;
  }
  __syncthreads();
}

// coarsest level: <= cmaxit x [lexicographic GS sweep along anti-diagonals, residual, h-scaled norm], strict arithmetic
template <typename TC>
__device__ void coarse_solve(TC* u, const TC* f, int nx, int ny, const StencilScalars<TC>& s, double hxhy, double tol,
                             int maxit, double* red, double* info) {
  int it = 1;
  double norm = 0.0;
  if (nx - 2 <= 32 && ny - 2 <= 32 && nx * ny <= 1024) {
    // tiny grid: one warp does everything with warp-level synchronisation; the rest of the block just waits
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      for (; it <= maxit; ++it) {
        for (int d = 2; d <= nx + ny - 4; ++d) {
          const int ilo = max(1, d - (ny - 2)), ihi = min(nx - 2, d - 1);
          const int i = ilo + lane;
          if (i <= ihi) {
            const int j = d - i;
            TC* p = u + i * ny + j;
            p[0] = relax_strict<TC>(s, p[0], p[ny], p[-ny], p[1], p[-1], f[i * ny + j]);
          }
          __syncwarp();
        }
        double acc = 0.0;
        for (int k = lane; k < nx * ny; k += 32) {
          const int i = k / ny, j = k - i * ny;
          TC v = f[k];
          if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
            const TC* p = u + k;
            v = Strict<TC>::sub(v, apply_strict<TC>(s, p[0], p[ny], p[-ny], p[1], p[-1]));
          }
          acc += (double)Strict<TC>::mul(v, v);
        }
        acc = warp_sum(acc);
        norm = sqrt(hxhy * acc);
        __syncwarp();
        if (norm < tol) break;
      }
      if (info != nullptr && lane == 0) {
        info[0] = (double)(it > maxit ? maxit : it);
        info[1] = norm;
      }
    }
    __syncthreads();
    return;
  }
  for (; it <= maxit; ++it) {
    for (int d = 2; d <= nx + ny - 4; ++d) {
      const int ilo = max(1, d - (ny - 2)), ihi = min(nx - 2, d - 1);
      for (int i = ilo + (int)threadIdx.x; i <= ihi; i += THREADS) {
        const int j = d - i;
        TC* p = u + i * ny + j;
        p[0] = relax_strict<TC>(s, p[0], p[ny], p[-ny], p[1], p[-1], f[i * ny + j]);
      }
      __syncthreads();
    }
    double acc = 0.0;
    for (int k = threadIdx.x; k < nx * ny; k += THREADS) {
      const int i = k / ny, j = k - i * ny;
      TC v = f[k];
      if (i > 0 && i < nx - 1 && j > 0 && j < ny - 1) {
        const TC* p = u + k;
        v = Strict<TC>::sub(v, apply_strict<TC>(s, p[0], p[ny], p[-ny], p[1], p[-1]));
      }
      acc += (double)Strict<TC>::mul(v, v);
    }
    acc = block_reduce(acc, red);
    norm = sqrt(hxhy * acc);
    if (norm < tol) break;
  }
  if (info != nullptr && threadIdx.x == 0) {
    info[0] = (double)(it > maxit ? maxit : it);
    info[1] = norm;
  }
  __syncthreads();
}

template <typename T, typename TC>
__global__ void __launch_bounds__(THREADS) small_cycle_kernel(const Params<T, TC> p) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ double red[32];
  const int L = p.nlev, last = L - 1;
  auto U = [&](int l) { return reinterpret_cast<T*>(sm + p.off_u[l]); };
  auto F = [&](int l) { return reinterpret_cast<T*>(sm + p.off_f[l]); };
  TC* const uc = reinterpret_cast<TC*>(sm + p.off_u[last]);
  TC* const fc = reinterpret_cast<TC*>(sm + p.off_f[last]);

  // entry level: global -> shared
  {
    const int nx = p.nx[0], ny = p.ny[0];
    for (int k = threadIdx.x; k < nx * ny; k += THREADS) {
      const int i = k / ny, j = k - i * ny;
      const T fv = p.f[(int64_t)i * p.ld_f + j];
      const T uv = p.u_zero ? (T)0 : p.u[(int64_t)i * p.ld_u + j];
      if (L == 1) { uc[k] = (TC)uv; fc[k] = (TC)fv; }
      else { U(0)[k] = uv; F(0)[k] = fv; }
    }
    __syncthreads();
  }

  if (L == 1) {
    coarse_solve<TC>(uc, fc, p.nx[0], p.ny[0], p.scc, p.hxhy_c, p.ctol, p.cmaxit, red, p.info);
  } else {
    int rep[MAXLEV];
    int l = 0;
    bool down = true;
    while (true) {
      if (l == last) {
        coarse_solve<TC>(uc, fc, p.nx[l], p.ny[l], p.scc, p.hxhy_c, p.ctol, p.cmaxit, red, p.info);
        l -= 1;
        down = false;
        continue;
      }
      if (down) {
        smooth<T>(U(l), F(l), p.nx[l], p.ny[l], p.sc[l], p.pre, p.iso1 != 0);
        const int nxc = p.nx[l + 1], nyc = p.ny[l + 1];
        if (l + 1 == last) {
          restrict_residual<T, TC>(U(l), F(l), p.nx[l], p.ny[l], p.sc[l], fc, nxc, nyc, p.iso1 != 0);
          for (int k = threadIdx.x; k < nxc * nyc; k += THREADS) uc[k] = (TC)0;
        } else {
          restrict_residual<T, T>(U(l), F(l), p.nx[l], p.ny[l], p.sc[l], F(l + 1), nxc, nyc, p.iso1 != 0);
          for (int k = threadIdx.x; k < nxc * nyc; k += THREADS) U(l + 1)[k] = (T)0;
        }
        __syncthreads();
        rep[l] = 0;
        l += 1;
      } else {  // a child cycle on level l+1 has just finished
        rep[l] += 1;
        const int reps = p.cycle == 0 ? 1 : (p.cycle == 1 ? 2 : max(1, 1 << max(0, L - l - 2)));
        if (rep[l] < reps) {
          l += 1;
          down = true;
          continue;
        }
        if (l + 1 == last) prolong_add<T, TC>(U(l), p.nx[l], p.ny[l], uc, p.ny[l + 1]);
        else prolong_add<T, T>(U(l), p.nx[l], p.ny[l], U(l + 1), p.ny[l + 1]);
        smooth<T>(U(l), F(l), p.nx[l], p.ny[l], p.sc[l], p.post, p.iso1 != 0);
        if (l == 0) break;
        l -= 1;
      }
    }
  }

  // entry level: shared -> global (boundary included: the correction may have touched it, multigrid.py:329)
  {
    const int nx = p.nx[0], ny = p.ny[0];
    for (int k = threadIdx.x; k < nx * ny; k += THREADS) {
      const int i = k / ny, j = k - i * ny;
      p.u[(int64_t)i * p.ld_u + j] = (L == 1) ? (T)uc[k] : U(0)[k];
    }
  }
}

template <typename T, typename TC>
static int launch(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy, int nlev,
                  int cycle, int pre, int post, double omega, double coefficient, double shift, double ctol, int cmaxit,
                  int u_zero, double* info, cudaStream_t st) {
  Params<T, TC> p;
  memset(&p, 0, sizeof(p));
  if (nlev < 1 || nlev > MAXLEV) return MG_ERR_UNSUPPORTED;
  size_t off = 0;
  int n = nx, m = ny;
  double hxl = hx, hyl = hy;
  for (int l = 0; l < nlev; ++l) {
    if (n < 3 || m < 3) return MG_ERR_BADARG;
    p.nx[l] = n; p.ny[l] = m;
    const size_t esz = (l == nlev - 1) ? sizeof(TC) : sizeof(T);
    const size_t bytes = ((size_t)n * m * esz + 15) & ~(size_t)15;
    p.off_u[l] = (unsigned)off; off += bytes;
    p.off_f[l] = (unsigned)off; off += bytes;
    if (l < nlev - 1) {
      p.sc[l] = make_scalars<T>(hxl, hyl, omega, coefficient, shift);
      if ((n - 1) % 2 || (m - 1) % 2) return MG_ERR_BADARG;
      n = (n - 1) / 2 + 1; m = (m - 1) / 2 + 1; hxl *= 2; hyl *= 2;
    } else {
      p.scc = make_scalars<TC>(hxl, hyl, 1.0, coefficient, shift);  // the coarse solver is plain GS (omega = 1)
      p.hxhy_c = hxl * hyl;
    }
  }
  if (off > 200 * 1024) return MG_ERR_UNSUPPORTED;
  p.nlev = nlev; p.cycle = cycle; p.pre = pre; p.post = post; p.u_zero = u_zero;
  p.iso1 = (omega == 1.0 && hx == hy) ? 1 : 0;  // the same selection as mg_stream_api.cu
  p.ctol = ctol; p.cmaxit = cmaxit; p.u = (T*)u; p.f = (const T*)f; p.ld_u = ld_u; p.ld_f = ld_f; p.info = info;
  auto kern = small_cycle_kernel<T, TC>;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured[dev] = true;
  }
  kern<<<1, THREADS, off, st>>>(p);
  return MG_OK;
}

}  // namespace small
}  // namespace mg

using namespace mg;

extern "C" {

int mg_small_cycle_smem_bytes(int nx, int ny, int nlev, int dtype, int coarse_dtype) {
  size_t off = 0;
  int n = nx, m = ny;
  for (int l = 0; l < nlev; ++l) {
    if (n < 3 || m < 3) return -1;
    const size_t esz = (l == nlev - 1) ? (coarse_dtype == MG_F64 ? 8 : 4) : (dtype == MG_F64 ? 8 : 4);
    off += 2 * (((size_t)n * m * esz + 15) & ~(size_t)15);
    if (l < nlev - 1) {
      if ((n - 1) % 2 || (m - 1) % 2) return -1;
      n = (n - 1) / 2 + 1; m = (m - 1) / 2 + 1;
    }
  }
  return off > (size_t)INT32_MAX ? -1 : (int)off;
}

int mg_small_cycle(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy, int nlev,
                   int cycle, int pre, int post, double omega, double coefficient, double shift, double coarse_tolerance,
                   int coarse_max_iterations, int u_zero, double* info, int dtype, int coarse_dtype, void* stream) {
  if (!u || !f || nx < 3 || ny < 3 || ld_u < ny || ld_f < ny || hx <= 0 || hy <= 0 || pre < 0 || post < 0 ||
      cycle < 0 || cycle > 2 || coarse_max_iterations < 1 || !(shift >= 0))
    return MG_ERR_BADARG;
  cudaStream_t st = as_stream(stream);
  int rc;
  if (dtype == MG_F64 && coarse_dtype == MG_F64)
    rc = small::launch<double, double>(u, f, nx, ny, ld_u, ld_f, hx, hy, nlev, cycle, pre, post, omega, coefficient, shift,
                                       coarse_tolerance, coarse_max_iterations, u_zero, info, st);
  else if (dtype == MG_F32 && coarse_dtype == MG_F64)
    rc = small::launch<float, double>(u, f, nx, ny, ld_u, ld_f, hx, hy, nlev, cycle, pre, post, omega, coefficient, shift,
                                      coarse_tolerance, coarse_max_iterations, u_zero, info, st);
  else if (dtype == MG_F32 && coarse_dtype == MG_F32)
    rc = small::launch<float, float>(u, f, nx, ny, ld_u, ld_f, hx, hy, nlev, cycle, pre, post, omega, coefficient, shift,
                                     coarse_tolerance, coarse_max_iterations, u_zero, info, st);
  else
    return MG_ERR_DTYPE;
  if (rc != MG_OK) return rc;
  return check_launch("mg_small_cycle");
}

}  // extern "C"
