// libmgb200: kernels of the reference's SECONDARY solver, CorrectedMultigridSolver
// (src/multigrid/solvers/corrected_multigrid.py:24-418) -- the V-cycle its validation modules and tutorials run.
// Its arithmetic differs from the primary path on purpose (see oracle/corrected_oracle.py): a Gauss-Seidel point
// update written as 0.25 * (W + E + S + N + h^2 f), a residual f - (-lap_h u) with a zero ring, full weighting / 16 on
// interior coarse points only, textbook bilinear prolongation, unscaled Frobenius norms.  Every expression below
// keeps the reference's operand order with IEEE round-to-nearest operations and no fused multiply-add (Strict<T>), so
// solutions are the reference's bit for bit; norms agree up to the order of the summation tree.
// fp64 only, like the reference class (its arrays are NumPy float64 whatever the precision manager says).
#include "mg_common.cuh"
#include "mg_lexgs.cuh"
#include "../../include/mgb200.h"

namespace mg {
namespace corrected {

using A = Strict<double>;
constexpr int CX = 32, CY = 8;       // 2-D tile of the pointwise kernels
constexpr int RBLOCKS = 512, RTHREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed tree: lanes by shuffle, warps through shared memory in warp order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(RTHREADS) finish_sum_kernel(const double* partials, int n, double* out) {
  __shared__ double sh[RTHREADS / 32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += RTHREADS) v += partials[i];
  const double t = block_sum(v, sh);
  if (threadIdx.x == 0) out[0] = t;
}

// r = f - (-(u[i-1,j] + u[i+1,j] + u[i,j-1] + u[i,j+1] - 4 u) / h^2) inside, 0 on the ring (:279-308);
// partials[block] = sum of r^2 over the rows the block owns (rows blockIdx.x, blockIdx.x + gridDim.x, ...)
__global__ void __launch_bounds__(RTHREADS) residual_kernel(const double* __restrict__ u, const double* __restrict__ f,
                                                            double* __restrict__ r, double* partials, int nx, int ny,
                                                            int64_t ldu, int64_t ldf, int64_t ldr, double h2) {
  __shared__ double sh[RTHREADS / 32];
  double acc = 0.0;
  for (int i = blockIdx.x; i < nx; i += gridDim.x) {
    const bool brow = i == 0 || i == nx - 1;
    for (int j = threadIdx.x; j < ny; j += RTHREADS) {
      double v = 0.0;
      if (!brow && j > 0 && j < ny - 1) {
        const double* p = u + (int64_t)i * ldu + j;
        const double t = A::sub(A::add(A::add(A::add(p[-ldu], p[ldu]), p[-1]), p[1]), A::mul(4.0, p[0]));
        v = A::sub(f[(int64_t)i * ldf + j], A::div(-t, h2));
      }
      if (r) r[(int64_t)i * ldr + j] = v;
      acc = A::add(acc, A::mul(v, v));
    }
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0 && partials) partials[blockIdx.x] = t;
}

// partials[block] = sum (a - b)^2  (||u - u_old|| of the coarsest-grid iteration, :384-387)
__global__ void __launch_bounds__(RTHREADS) diff_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                        double* partials, int nx, int ny, int64_t lda, int64_t ldb) {
  __shared__ double sh[RTHREADS / 32];
  double acc = 0.0;
  for (int i = blockIdx.x; i < nx; i += gridDim.x)
    for (int j = threadIdx.x; j < ny; j += RTHREADS) {
      const double d = A::sub(a[(int64_t)i * lda + j], b[(int64_t)i * ldb + j]);
      acc = A::add(acc, A::mul(d, d));
    }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}

// Full weighting on the interior coarse points whose 3x3 fine neighbourhood exists, 0 elsewhere (:318-335)
__global__ void __launch_bounds__(CX* CY) restrict_kernel(const double* __restrict__ fine, double* __restrict__ coarse,
                                                           int nxf, int nyf, int nxc, int nyc, int64_t ldf, int64_t ldc) {
  const int J = blockIdx.x * CX + threadIdx.x, I = blockIdx.y * CY + threadIdx.y;
  if (I >= nxc || J >= nyc) return;
  double v = 0.0;
  if (I >= 1 && I <= nxc - 2 && J >= 1 && J <= nyc - 2 && 2 * I < nxf - 1 && 2 * J < nyf - 1) {
    const double* p = fine + (int64_t)(2 * I) * ldf + 2 * J;
    double s = A::add(p[-ldf - 1], A::mul(2.0, p[-ldf]));
    s = A::add(s, p[-ldf + 1]);
    s = A::add(s, A::mul(2.0, p[-1]));
    s = A::add(s, A::mul(4.0, p[0]));
    s = A::add(s, A::mul(2.0, p[1]));
    s = A::add(s, p[ldf - 1]);
    s = A::add(s, A::mul(2.0, p[ldf]));
    s = A::add(s, p[ldf + 1]);
    v = A::div(s, 16.0);
  }
  coarse[(int64_t)I * ldc + J] = v;
}

// u <- u + P(coarse) with the textbook bilinear prolongation (:337-364), then the ring of u zeroed (:231-233)
__global__ void __launch_bounds__(CX* CY) prolong_add_kernel(const double* __restrict__ c, double* __restrict__ u, int nxc,
                                                              int nyc, int nxf, int nyf, int64_t ldc, int64_t ldu) {
  const int j = blockIdx.x * CX + threadIdx.x, i = blockIdx.y * CY + threadIdx.y;
  if (i >= nxf || j >= nyf) return;
  double* q = u + (int64_t)i * ldu + j;
  if (i == 0 || j == 0 || i == nxf - 1 || j == nyf - 1) {
    q[0] = 0.0;
    return;
  }
  const int ic = i >> 1, jc = j >> 1;
  const bool oi = i & 1, oj = j & 1;
  double p = 0.0;
  if (ic < nxc && jc < nyc && (!oi || ic + 1 < nxc) && (!oj || jc + 1 < nyc)) {
    const double* b = c + (int64_t)ic * ldc + jc;
    if (!oi && !oj) p = b[0];
    else if (!oi) p = A::mul(0.5, A::add(b[0], b[1]));
    else if (!oj) p = A::mul(0.5, A::add(b[0], b[ldc]));
    else p = A::mul(0.25, A::add(A::add(A::add(b[0], b[ldc]), b[1]), b[ldc + 1]));
  }
  q[0] = A::add(q[0], p);
}

}  // namespace corrected
}  // namespace mg

using namespace mg;
using namespace mg::corrected;

extern "C" {

int mg_cm_workspace_doubles(void) { return RBLOCKS; }

int mg_cm_gs(double* u, const double* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double h, int sweeps, void* stream) {
  MG_REQUIRE(u && f && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && h > 0 && sweeps >= 0);
  const int nwarps = (nx - 2 + 31) / 32;
  const auto sc = make_scalars<double>(h, h, 1.0, 1.0);  // only hx2 = h ** 2 is used
  cudaStream_t st = as_stream(stream);
  for (int k = 0; k < sweeps; ++k) {
    int* prog = lexgs_progress(nwarps);
    if (!prog) return MG_ERR_UNSUPPORTED;
    if (cudaMemsetAsync(prog, 0, sizeof(int) * nwarps, st) != cudaSuccess) return MG_ERR_LAUNCH;
    lexgs_pipe_kernel<double, true, true><<<nwarps, 32, 0, st>>>(u, f, nx, ny, ld_u, ld_f, prog, sc);
  }
  return check_launch("mg_cm_gs", sweeps);
}

int mg_cm_residual(const double* u, const double* f, double* r, double* sumsq_out, double* workspace, int nx, int ny,
                   int64_t ld_u, int64_t ld_f, int64_t ld_r, double h, void* stream) {
  MG_REQUIRE(u && f && nx >= 3 && ny >= 3 && ld_u >= ny && ld_f >= ny && h > 0 && (!r || ld_r >= ny));
  MG_REQUIRE((sumsq_out == nullptr) == (workspace == nullptr) && (r || sumsq_out));
  cudaStream_t st = as_stream(stream);
  const int blocks = nx < RBLOCKS ? nx : RBLOCKS;
  residual_kernel<<<blocks, RTHREADS, 0, st>>>(u, f, r, workspace, nx, ny, ld_u, ld_f, ld_r, pow(h, 2.0));
  if (sumsq_out) finish_sum_kernel<<<1, RTHREADS, 0, st>>>(workspace, blocks, sumsq_out);
  return check_launch("mg_cm_residual", sumsq_out ? 2 : 1);
}

int mg_cm_diff_sumsq(const double* a, const double* b, double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_a,
                     int64_t ld_b, void* stream) {
  MG_REQUIRE(a && b && sumsq_out && workspace && nx >= 1 && ny >= 1 && ld_a >= ny && ld_b >= ny);
  cudaStream_t st = as_stream(stream);
  const int blocks = nx < RBLOCKS ? nx : RBLOCKS;
  diff_kernel<<<blocks, RTHREADS, 0, st>>>(a, b, workspace, nx, ny, ld_a, ld_b);
  finish_sum_kernel<<<1, RTHREADS, 0, st>>>(workspace, blocks, sumsq_out);
  return check_launch("mg_cm_diff_sumsq", 2);
}

int mg_cm_restrict(const double* fine, double* coarse, int nxf, int nyf, int nxc, int nyc, int64_t ld_f, int64_t ld_c,
                   void* stream) {
  MG_REQUIRE(fine && coarse && nxf >= 3 && nyf >= 3 && nxc >= 3 && nyc >= 3 && ld_f >= nyf && ld_c >= nyc);
  const dim3 g((nyc + CX - 1) / CX, (nxc + CY - 1) / CY), b(CX, CY);
  restrict_kernel<<<g, b, 0, as_stream(stream)>>>(fine, coarse, nxf, nyf, nxc, nyc, ld_f, ld_c);
  return check_launch("mg_cm_restrict");
}

int mg_cm_prolong_add(const double* coarse, double* u, int nxc, int nyc, int nxf, int nyf, int64_t ld_c, int64_t ld_u,
                      void* stream) {
  MG_REQUIRE(coarse && u && nxf >= 3 && nyf >= 3 && nxc >= 2 && nyc >= 2 && ld_u >= nyf && ld_c >= nyc);
  const dim3 g((nyf + CX - 1) / CX, (nxf + CY - 1) / CY), b(CX, CY);
  prolong_add_kernel<<<g, b, 0, as_stream(stream)>>>(coarse, u, nxc, nyc, nxf, nyf, ld_c, ld_u);
  return check_launch("mg_cm_prolong_add");
}

}  // extern "C"
