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
constexpr int THREADS = 512;  // 128 registers per thread: the register-resident coarsest solve must not spill

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
  int profile;  // 1: info[2..9] += SM cycles spent per phase kind (see mg_small_cycle in mgb200.h); info holds 16 doubles
  int iso1;  // hx == hy and omega == 1: the streaming kernel's 5-instruction point update (same bits as there)
  int exact5;      // coarsest grid is 5 x 5 with dyadic isotropic spacing, no shift, coefficient -1: register solver
  double xthr;     // largest double whose square root is below ctol (exact5 stopping test without a sqrt per sweep)
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

// ---------------------------------------------------------------------------------------------------------
// Phases.  Every level is worked on by the whole block: rows over warps, columns over lanes, one barrier per phase.
// Measured on B200 (profiles/r02_small_cycle_profile.json) and kept in mind below:
//   * on the tiny levels a phase is a handful of instructions per thread, executed ONCE per visit, so its cost is the
//     instruction fetch of cold straight-line code (~1000 cycles per phase when the body was unrolled four-fold with
//     grouped loads, ~3x less as a compact loop): the phase functions are small noinline functions with rolled loops,
//     specialised at compile time (ISO1) instead of carrying both point updates;
//   * handing the 9 x 9 and 17 x 17 levels to ONE warp (no block barrier) lost: a single warp walks 2-4 row groups per
//     half-sweep where 16 warps take one row each; only the 5 x 5 coarsest solve runs in one thread;
//   * on the large levels (129^2, 65^2) a half-sweep is bound by shared-memory bandwidth (stride-2 red-black accesses
//     use half of every 128-byte wavefront).
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool ISO1>
__device__ __noinline__ void smooth(T* u, const T* f, int nx, int ny, const StencilScalars<T>* sp, int sweeps) {
  const StencilScalars<T> s = *sp;  // shared memory -> registers (only the fields the chosen update reads)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int h = 0; h < 2 * sweeps; ++h) {  // half-sweeps: red (row + column even) first
#pragma unroll 1
    for (int i = 1 + warp; i <= nx - 2; i += THREADS / 32) {
      T* row = u + i * ny;
      const T* frow = f + i * ny;
#pragma unroll 1
      for (int j = 1 + ((i + 1 + h) & 1) + 2 * lane; j <= ny - 2; j += 64) {
        T* p = row + j;
        p[0] = ISO1 ? relax_iso1<T>(s, p[ny], p[-ny], p[1], p[-1], frow[j])
                    : relax_fast<T>(s, p[0], p[ny], p[-ny], p[1], p[-1], frow[j]);
      }
    }
    __syncthreads();
  }
}

// Full weighting of the residual at ONE interior coarse point (I, J) from the 5 x 5 patch of u around fine (2I, 2J): the
// nine residuals it averages are all at interior fine points (2I - 1 >= 1), so there is no boundary case and 21 + 9 loads
// replace the 9 x 6 of nine independent residual evaluations.  Reference summation order (transfer.py:116-122).
template <typename T, bool ISO1>
__device__ __forceinline__ T fw_residual_at(const T* u, const T* f, int ny, int i, int j, const StencilScalars<T>& s) {
  T w[5][5];
#pragma unroll
  for (int a = 0; a < 5; ++a)
#pragma unroll
    for (int b = 0; b < 5; ++b)
      if (!((a == 0 || a == 4) && (b == 0 || b == 4))) w[a][b] = u[(i - 2 + a) * ny + (j - 2 + b)];
  T r[3][3];
#pragma unroll
  for (int a = 1; a <= 3; ++a)
#pragma unroll
    for (int b = 1; b <= 3; ++b) {
      const T fv = f[(i - 2 + a) * ny + (j - 2 + b)];
      r[a - 1][b - 1] = ISO1 ? residual_iso<T>(s, w[a][b], w[a + 1][b], w[a - 1][b], w[a][b + 1], w[a][b - 1], fv)
                             : residual_fast<T>(s, w[a][b], w[a + 1][b], w[a - 1][b], w[a][b + 1], w[a][b - 1], fv);
    }
  const T corners = ((r[0][0] + r[0][2]) + r[2][0]) + r[2][2];
  const T edges = ((r[0][1] + r[2][1]) + r[1][0]) + r[1][2];
  return ((T)0.0625 * corners + (T)0.125 * edges) + (T)0.25 * r[1][1];
}

// f_c = R(f - A u): injection on the coarse boundary, full weighting inside; then the coarse iterate is zeroed
template <typename T, typename TO, bool ISO1>
__device__ __noinline__ void restrict_residual(const T* u, const T* f, int nx, int ny, const StencilScalars<T>* sp, TO* fc,
                                               TO* uc, int nxc, int nyc) {
  const StencilScalars<T> s = *sp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int I = warp; I < nxc; I += THREADS / 32)
#pragma unroll 1
    for (int J = lane; J < nyc; J += 32) {
      const int i = 2 * I, j = 2 * J;
      T v;
      if (I == 0 || I == nxc - 1 || J == 0 || J == nyc - 1) v = resid_at<T>(u, f, nx, ny, i, j, s, ISO1);
      else v = fw_residual_at<T, ISO1>(u, f, ny, i, j, s);
      fc[I * nyc + J] = (TO)v;
      uc[I * nyc + J] = (TO)0;  // the coarse error equation starts from zero (multigrid.py:303)
    }
  __syncthreads();
}

// u += P e_c, bilinear with the reference's last-row / last-column treatment (transfer.py:234-267)
template <typename T, typename TI>
__device__ __noinline__ void prolong_add(T* u, int nx, int ny, const TI* ec, int nyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int i = warp; i < nx; i += THREADS / 32)
#pragma unroll 1
    for (int j = lane; j < ny; j += 32) {
      const TI* c = ec + (i >> 1) * nyc + (j >> 1);
      const bool oi = i & 1, oj = j & 1;
      // all four coarse neighbours, clamped into the array (unused ones drop out below): no divergent loads
      const T c00 = (T)c[0], c01 = (T)c[oj ? 1 : 0], c10 = (T)c[oi ? nyc : 0], c11 = (T)c[(oi ? nyc : 0) + (oj ? 1 : 0)];
      T v;
      if (!oi && !oj) v = c00;
      else if (oi && !oj) v = (j < ny - 1) ? (T)0.5 * (c00 + c10) : (T)0;
      else if (!oi && oj) v = (i < nx - 1) ? (T)0.5 * (c00 + c01) : (T)0;
      else v = (T)0.25 * (((c00 + c01) + c10) + c11);
      u[i * ny + j] += v;
    }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// Coarsest level: <= cmaxit x [lexicographic GS sweep, residual, h-scaled norm] (solvers/base.py:258-285).
// Every variant evaluates the SAME rounded operations in the same order as mg_coarse_solve_lexgs.
// ---------------------------------------------------------------------------------------------------------

// (a) The 5 x 5 coarsest grid of every 2^k + 1 hierarchy, isotropic dyadic spacing, omega = 1, no shift, coefficient -1:
// ONE THREAD, all 25 values in registers.  A W-cycle over 14 levels makes 8192 coarsest solves of ~12 sweeps each, all
// on the critical path, so what matters is the latency of one sweep + stopping test:
//   * power-of-two scalings commute with rounding, so the point update ((up+dn)/h^2 + (rt+lf)/h^2 + f) / (4/h^2) equals
//     (((up+dn)+(rt+lf)) + h^2 f) * 0.25 and the residual f - (-(S/h^2 - u*4/h^2)) equals (h^2 f + fma(-4, u, S)) / h^2
//     bit for bit (checked on the CPU against the strict expressions on 12M random inputs); 4 and 3 dependent
//     operations instead of 9 and 9;
//   * the squared residuals are summed in the order of the strict kernel's lane-strided warp tree (leaf k = point k,
//     partners k ^ 16, 8, 4, 2, 1); the common factor 1/h^4 is applied once to the sum (exact);
//   * sqrt is monotone and correctly rounded, so "sqrt(x) < tol" is decided as "x <= xthr" with xthr the largest double
//     whose root is below tol (found on the host);
//   * the test of sweep k is evaluated while sweep k+1 is already running (software pipelining: both are straight-line
//     code in one basic block; there is no speculation in hardware), and the extra sweep is dropped when k passed.
template <typename TC> struct Scal5 { TC hx2, ihx2, inv_neg_diag; };  // by value: no local copy of the kernel parameters

// ZB: the boundary values of u and of f are all zero (every coarse error equation of a solve whose right-hand side
// vanishes on the boundary ring -- the facade's default): the boundary terms drop out of the sums (x + 0 = x) and out
// of the norm tree, which leaves 27 + 44 fp64 instructions per sweep + test and few enough live values for registers.
template <typename TC, bool ZB>
__device__ __forceinline__ void coarse_solve_5x5_exact(TC* u, const TC* f, const Scal5<TC>& s, double hxhy,
                                                       double xthr, int maxit, double* info) {
  using A = Strict<TC>;
  TC U[5][5], G[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const bool in = i >= 1 && i <= 3 && j >= 1 && j <= 3;
      U[i][j] = (ZB && !in) ? (TC)0 : u[i * 5 + j];
      G[i][j] = (ZB && !in) ? (TC)0 : A::mul(f[i * 5 + j], s.hx2);  // h^2 f (exact scaling)
    }
  const TC quarter = A::mul(s.ihx2, s.inv_neg_diag);       // 1/h^2 * h^2/4 = 1/4 exactly
  const double ih4 = (double)s.ihx2 * (double)s.ihx2;      // power of two

  // x + y where an operand known to be the zero boundary value is dropped at compile time (after unrolling)
  auto add2 = [&](bool ha, TC a, bool hb, TC b) -> TC {
    if (ha && hb) return A::add(a, b);
    return ha ? a : (hb ? b : (TC)0);
  };
  auto nbsum = [&](const TC(&V)[5][5], int i, int j) -> TC {
    const bool hu = !ZB || i + 1 <= 3, hd = !ZB || i - 1 >= 1, hr = !ZB || j + 1 <= 3, hl = !ZB || j - 1 >= 1;
    const TC t1 = add2(hu, V[i + 1][j], hd, V[i - 1][j]);
    const TC t2 = add2(hr, V[i][j + 1], hl, V[i][j - 1]);
    return add2(hu || hd, t1, hr || hl, t2);
  };
  auto sweep = [&](TC(&V)[5][5]) {
#pragma unroll
    for (int i = 1; i <= 3; ++i)
#pragma unroll
      for (int j = 1; j <= 3; ++j) V[i][j] = A::mul(A::add(G[i][j], nbsum(V, i, j)), quarter);
  };
  auto normx = [&](const TC(&V)[5][5]) -> double {
    // leaves of the strict kernel's warp tree: leaf k = point k = 5 i + j, partners k ^ 16, 8, 4, 2, 1
    double a[32];
    bool nz[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { a[k] = 0.0; nz[k] = false; }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const bool in = i >= 1 && i <= 3 && j >= 1 && j <= 3;
        if (in) {
          const TC w = A::add(G[i][j], fma((TC)-4, V[i][j], nbsum(V, i, j)));   // h^2 (f - A u)
          a[i * 5 + j] = (double)A::mul(w, w);
          nz[i * 5 + j] = true;
        } else if (!ZB) {
          a[i * 5 + j] = (double)A::mul(G[i][j], G[i][j]);  // r = f on the boundary (loop invariant: hoisted)
          nz[i * 5 + j] = true;
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < o; ++k) {
        if (nz[k] && nz[k + o]) a[k] = __dadd_rn(a[k], a[k + o]);
        else if (nz[k + o]) a[k] = a[k + o];                 // 0 + x = x for the non-negative squares
        nz[k] = nz[k] || nz[k + o];
      }
    return __dmul_rn(hxhy, __dmul_rn(a[0], ih4));
  };

  TC W[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 5; ++j) W[i][j] = U[i][j];
  sweep(U);  // U = state after sweep 1
  int it = 1;
  double x;
  while (true) {
#pragma unroll
    for (int i = 1; i <= 3; ++i)
#pragma unroll
      for (int j = 1; j <= 3; ++j) W[i][j] = U[i][j];
    sweep(W);         // state after sweep it + 1, started before the test of sweep `it` is known
    x = normx(U);
    if (x <= xthr || it >= maxit) break;
#pragma unroll
    for (int i = 1; i <= 3; ++i)
#pragma unroll
      for (int j = 1; j <= 3; ++j) U[i][j] = W[i][j];
    ++it;
  }
#pragma unroll
  for (int i = 1; i <= 3; ++i)
#pragma unroll
    for (int j = 1; j <= 3; ++j) u[i * 5 + j] = U[i][j];
  if (info != nullptr) {
    info[0] = (double)it;
    info[1] = sqrt(x);
  }
}

// zb: the boundary rings of the coarsest u and f are zero.  Decided ONCE per launch from the entry level (coarse
// boundary values of f are injected from the entry ring, transfer.py:109-113; coarse iterates start from zero), not
// per call: the prologue of a coarsest solve is cold straight-line code and a W-cycle runs it thousands of times.
template <typename TC>
__device__ __noinline__ void coarse_solve_5x5(TC* u, const TC* f, const Scal5<TC> s, double hxhy, double xthr, int maxit,
                                              double* info, int zb) {
  if (zb) coarse_solve_5x5_exact<TC, true>(u, f, s, hxhy, xthr, maxit, info);
  else coarse_solve_5x5_exact<TC, false>(u, f, s, hxhy, xthr, maxit, info);
}

// (b) Any coarsest grid of at most 32 points: one point per LANE, values in registers, neighbours by warp shuffle (no
// shared-memory round trip per anti-diagonal).  Strict arithmetic, same sweep order, residual, summation tree and stopping
// rule as (c).
template <typename TC>
__device__ __forceinline__ void coarse_solve_lanes(TC* u, const TC* f, int nx, int ny, const StencilScalars<TC>& s, double hxhy, double tol,
                                   int maxit, double* info) {
  const int lane = threadIdx.x & 31, n = nx * ny;
  const int k = lane < n ? lane : 0;
  const int i = k / ny, j = k - i * ny;
  const bool inside = lane < n && i > 0 && i < nx - 1 && j > 0 && j < ny - 1;
  TC uk = lane < n ? u[k] : (TC)0;
  const TC fk = lane < n ? f[k] : (TC)0;
  // shuffle sources, clamped into the warp (values fetched by non-interior lanes are never used)
  const int kup = min(k + ny, 31), kdn = max(k - ny, 0), krt = min(k + 1, 31), klf = max(k - 1, 0);
  int it = 1;
  double norm = 0.0;
  for (; it <= maxit; ++it) {
    for (int d = 2; d <= nx + ny - 4; ++d) {
      const TC up = __shfl_sync(0xffffffffu, uk, kup), dn = __shfl_sync(0xffffffffu, uk, kdn);
      const TC rt = __shfl_sync(0xffffffffu, uk, krt), lf = __shfl_sync(0xffffffffu, uk, klf);
      if (inside && i + j == d) uk = relax_strict<TC>(s, uk, up, dn, rt, lf, fk);
    }
    const TC up = __shfl_sync(0xffffffffu, uk, kup), dn = __shfl_sync(0xffffffffu, uk, kdn);
    const TC rt = __shfl_sync(0xffffffffu, uk, krt), lf = __shfl_sync(0xffffffffu, uk, klf);
    TC v = fk;  // r = f on the boundary; lanes beyond the grid hold f = 0
    if (inside) v = Strict<TC>::sub(v, apply_strict<TC>(s, uk, up, dn, rt, lf));
    double acc = 0.0;
    acc += (double)Strict<TC>::mul(v, v);
    acc = warp_sum(acc);
    norm = sqrt(hxhy * acc);
    if (norm < tol) break;
  }
  if (inside) u[k] = uk;
  if (info != nullptr && lane == 0) {
    info[0] = (double)(it > maxit ? maxit : it);
    info[1] = norm;
  }
}

// warp 0 only; no block barrier (the caller synchronises the team)
template <typename TC>
__device__ __forceinline__ void coarse_solve_warp(TC* u, const TC* f, int nx, int ny, const StencilScalars<TC>* sp, double hxhy,
                                               double tol, double xthr, int exact5, int maxit, double* info) {
  const StencilScalars<TC> s = *sp;
  const int lane = threadIdx.x & 31;
  if (exact5) {  // uniform; bit 1 = zero boundary rings
    if (lane == 0)
      coarse_solve_5x5<TC>(u, f, Scal5<TC>{s.hx2, s.ihx2, s.inv_neg_diag}, hxhy, xthr, maxit, info, exact5 & 2);
    __syncwarp();
    return;
  }
  if (nx * ny <= 32) {
    coarse_solve_lanes<TC>(u, f, nx, ny, s, hxhy, tol, maxit, info);
    __syncwarp();
    return;
  }
  // (c) up to 32 x 32 interior points: anti-diagonal wavefronts through shared memory
  int it = 1;
  double norm = 0.0;
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
  __syncwarp();
}

// whole block; ends with a block barrier
template <typename TC>
__device__ __noinline__ void coarse_solve(TC* u, const TC* f, int nx, int ny, const StencilScalars<TC>* sp, double hxhy,
                                          double tol, double xthr, int exact5, int maxit, double* red, double* info) {
  if (nx - 2 <= 32 && ny - 2 <= 32 && nx * ny <= 1024) {
    // tiny grid: one warp does everything with warp-level synchronisation; the rest of the block just waits
    if (threadIdx.x < 32) coarse_solve_warp<TC>(u, f, nx, ny, sp, hxhy, tol, xthr, exact5, maxit, info);
    __syncthreads();
    return;
  }
  const StencilScalars<TC> s = *sp;
  int it = 1;
  double norm = 0.0;
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

// The V / W / F recursion of solvers/multigrid.py:253-337 over all levels, iteratively.
template <typename T, typename TC, bool ISO1>
__device__ __forceinline__ void run_cycle(const Params<T, TC>& p, unsigned char* sm, double* red,
                                          const StencilScalars<T>* sc, const StencilScalars<TC>* scc, int exact5) {
  const int L = p.nlev, last = L - 1;
  auto U = [&](int l) { return reinterpret_cast<T*>(sm + p.off_u[l]); };
  auto F = [&](int l) { return reinterpret_cast<T*>(sm + p.off_f[l]); };
  TC* const uc = reinterpret_cast<TC*>(sm + p.off_u[last]);
  TC* const fc = reinterpret_cast<TC*>(sm + p.off_f[last]);
  int rep[MAXLEV];
  int l = 0;
  bool down = true;
  // optional phase profile (thread 0's clock; every phase ends with a block barrier, so this is the block's time)
  const bool prof = p.profile != 0 && p.info != nullptr && threadIdx.x == 0;
  long long t_prev = prof ? clock64() : 0;
  auto lap = [&](int slot) {
    if (prof) {
      const long long t = clock64();
      p.info[2 + slot] += (double)(t - t_prev);
      t_prev = t;
    }
  };
  while (true) {
    if (l == last) {
      coarse_solve<TC>(uc, fc, p.nx[l], p.ny[l], scc, p.hxhy_c, p.ctol, p.xthr, exact5, p.cmaxit, red, p.info);
      lap(3);
      if (l == 0) return;
      l -= 1;
      down = false;
      continue;
    }
    if (down) {
      smooth<T, ISO1>(U(l), F(l), p.nx[l], p.ny[l], sc + l, p.pre);
      lap(0);
      const int nxc = p.nx[l + 1], nyc = p.ny[l + 1];
      if (l + 1 == last) restrict_residual<T, TC, ISO1>(U(l), F(l), p.nx[l], p.ny[l], sc + l, fc, uc, nxc, nyc);
      else restrict_residual<T, T, ISO1>(U(l), F(l), p.nx[l], p.ny[l], sc + l, F(l + 1), U(l + 1), nxc, nyc);
      lap(1);
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
      lap(2);
      smooth<T, ISO1>(U(l), F(l), p.nx[l], p.ny[l], sc + l, p.post);
      lap(0);
      if (l == 0) return;
      l -= 1;
    }
  }
}

template <typename T, typename TC>
__global__ void __launch_bounds__(THREADS, 1) small_cycle_kernel(const Params<T, TC> p) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ double red[32];
  // per-level stencil scalars in shared memory: the phase functions are separate (noinline) functions with their own
  // register allocation and read them through a pointer (a reference into the kernel parameters would force a local copy)
  __shared__ StencilScalars<T> sc_s[MAXLEV];
  __shared__ StencilScalars<TC> scc_s;
  const int L = p.nlev, last = L - 1;
  auto U = [&](int l) { return reinterpret_cast<T*>(sm + p.off_u[l]); };
  auto F = [&](int l) { return reinterpret_cast<T*>(sm + p.off_f[l]); };
  TC* const uc = reinterpret_cast<TC*>(sm + p.off_u[last]);
  TC* const fc = reinterpret_cast<TC*>(sm + p.off_f[last]);
  if (threadIdx.x < MAXLEV) sc_s[threadIdx.x] = p.sc[threadIdx.x];
  if (threadIdx.x == 32) scc_s = p.scc;

  const long long t_begin = clock64();
  // entry level: global -> shared; on the way, are the boundary rings of u and f zero?
  int ring_zero;
  {
    const int nx = p.nx[0], ny = p.ny[0];
    int ok = 1;
    for (int k = threadIdx.x; k < nx * ny; k += THREADS) {
      const int i = k / ny, j = k - i * ny;
      const T fv = p.f[(int64_t)i * p.ld_f + j];
      const T uv = p.u_zero ? (T)0 : p.u[(int64_t)i * p.ld_u + j];
      if ((i == 0 || i == nx - 1 || j == 0 || j == ny - 1) && !(fv == (T)0 && uv == (T)0)) ok = 0;
      if (L == 1) { uc[k] = (TC)uv; fc[k] = (TC)fv; }
      else { U(0)[k] = uv; F(0)[k] = fv; }
    }
    ring_zero = __syncthreads_and(ok);
  }

  const long long t_loaded = clock64();
  const int exact5 = p.exact5 ? (1 | (ring_zero ? 2 : 0)) : 0;
  if (p.iso1) run_cycle<T, TC, true>(p, sm, red, sc_s, &scc_s, exact5);
  else run_cycle<T, TC, false>(p, sm, red, sc_s, &scc_s, exact5);
  __syncthreads();
  const long long t_cycled = clock64();

  // entry level: shared -> global (boundary included: the correction may have touched it, multigrid.py:329)
  {
    const int nx = p.nx[0], ny = p.ny[0];
    for (int k = threadIdx.x; k < nx * ny; k += THREADS) {
      const int i = k / ny, j = k - i * ny;
      p.u[(int64_t)i * p.ld_u + j] = (L == 1) ? (T)uc[k] : U(0)[k];
    }
  }
  if (p.profile != 0 && p.info != nullptr && threadIdx.x == 0) {
    p.info[10] += (double)(t_loaded - t_begin);
    p.info[11] += (double)(t_cycled - t_loaded);
    p.info[12] += (double)(clock64() - t_cycled);
    p.info[13] += 1.0;
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
  p.nlev = nlev; p.cycle = cycle; p.pre = pre; p.post = post; p.u_zero = u_zero & 1; p.profile = (u_zero >> 1) & 1;
  p.iso1 = (omega == 1.0 && hx == hy) ? 1 : 0;  // the same selection as mg_stream_api.cu
  p.exact5 = (p.nx[nlev - 1] == 5 && p.ny[nlev - 1] == 5 && p.scc.recip_exact && p.scc.hx2 == p.scc.hy2 &&
              p.scc.shift == (TC)0 && p.scc.coeff == (TC)-1) ? 1 : 0;
  p.xthr = -1.0;  // ctol <= 0: never
  if (ctol > 0) {
    double y = ctol * ctol;
    while (y > 0 && sqrt(y) >= ctol) y = nextafter(y, 0.0);
    while (sqrt(nextafter(y, INFINITY)) < ctol) y = nextafter(y, INFINITY);
    p.xthr = (sqrt(y) < ctol) ? y : -1.0;
  }
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
