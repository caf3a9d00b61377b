/*
 * mgb200.h  --  C ABI of libmgb200.so: the B200 (sm_100a) multigrid V/W-cycle hot path.
 *
 * The reference (Tani843/Mixed_Precision_Multigrid_Solvers_for_PDEs) has no FFI: its seam is
 * the Python operator protocol of src/multigrid/{operators,solvers}.  Every entry point below
 * names the reference method (file:line under /root/reference) whose arithmetic it replaces;
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Fields are row-major (nx, ny) arrays that INCLUDE the boundary points (core/grid.py:43-54);
 *     first index = x.  `ld` is the row pitch in ELEMENTS (>= ny).  Device pointers only.
 *   - dtype codes: MG_F32 = 0, MG_F64 = 1.
 *   - All calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default
 *     stream), never allocate, keep no global state and are thread-safe.
 *   - Return value: 0 = MG_OK, negative = error (see mg_status_string).  No exceptions cross
 *     the ABI.
 *   - "Vector path" kernels (the mg_vc_* family) need 16-byte aligned base pointers and
 *     ld % (16/sizeof(T)) == 0; they return MG_ERR_ALIGN otherwise.  The mg_* basic family has
 *     no alignment requirement.
 */
#ifndef MGB200_H
#define MGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MG_API __attribute__((visibility("default")))
#else
#define MG_API
#endif

#define MG_F32 0
#define MG_F64 1

#define MG_OK 0
#define MG_ERR_BADARG (-1)
#define MG_ERR_DTYPE (-2)
#define MG_ERR_ALIGN (-3)
#define MG_ERR_LAUNCH (-4)
#define MG_ERR_UNSUPPORTED (-5)

/* restriction / prolongation methods (operators/transfer.py:22-33, 158-169) */
#define MG_RESTRICT_FULL_WEIGHTING 0
#define MG_RESTRICT_INJECTION 1
#define MG_RESTRICT_HALF_WEIGHTING 2
#define MG_PROLONG_BILINEAR 0
#define MG_PROLONG_INJECTION 1

MG_API int mg_abi_version(void);
MG_API const char* mg_status_string(int status);
/* number of SMs of the current device, cached per device (used by callers to size workspaces) */
MG_API int mg_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Basic per-operator kernels (one reference method each)
 * ------------------------------------------------------------------------------------------- */

/* out = coefficient * lap_h(u) on the interior, 0 on the boundary.
 * Replaces LaplacianOperator.apply, operators/laplacian.py:44-80. */
MG_API int mg_apply_laplacian(const void* u, void* out, int nx, int ny, int64_t ld_u, int64_t ld_out,
                       double hx, double hy, double coefficient, int dtype, void* stream);

/* r = f - coefficient*lap_h(u) on the interior, r = f on the boundary.  u and f share dtype_in,
 * r may have a different dtype_out (fp32 u,f -> fp64 r is the reference's
 * mixed_precision_residual_kernel, gpu/cuda_kernels.py:843-883; the stencil is evaluated in
 * the wider of the two types).  Replaces LaplacianOperator.residual, operators/laplacian.py:105-124. */
MG_API int mg_residual(const void* u, const void* f, void* r, int nx, int ny, int64_t ld_u, int64_t ld_f,
                int64_t ld_r, double hx, double hy, double coefficient, int dtype_in, int dtype_out,
                void* stream);

/* In-place red-black Gauss-Seidel, `sweeps` x (red (i+j even) then black), relaxing
 * -lap_h(u) = f with over-relaxation omega.  Boundary never touched.
 * Replaces GaussSeidelSmoother(red_black=True).smooth, solvers/smoothers.py:117-151, 175-207
 * (and SmoothingKernels.red_black_gauss_seidel, gpu/cuda_kernels.py:348-390). */
MG_API int mg_smooth_rbgs(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx,
                   double hy, double omega, int sweeps, int dtype, void* stream);

/* Damped Jacobi, `sweeps` iterations ping-ponging between u and tmp (same shape/ld as u); the
 * result always ends in u.  Replaces JacobiSmoother.smooth / WeightedJacobiSmoother,
 * solvers/smoothers.py:41-86, 210-225. */
MG_API int mg_smooth_jacobi(void* u, void* tmp, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f,
                     double hx, double hy, double omega, int sweeps, int dtype, void* stream);

/* In-place lexicographic Gauss-Seidel (i-major, then j), evaluated along anti-diagonal
 * wavefronts by ONE thread block (identical arithmetic to the sequential sweep).
 * mode 0: forward sweeps  -- GaussSeidelSmoother(red_black=False).smooth, solvers/smoothers.py:153-173;
 * mode 1: backward sweeps (i, j descending)  -- _backward_sweep, solvers/smoothers.py:268-284;
 * mode 2: each sweep = forward then backward  -- SymmetricGaussSeidelSmoother.smooth, :246-266. */
#define MG_LEXGS_FORWARD 0
#define MG_LEXGS_BACKWARD 1
#define MG_LEXGS_SYMMETRIC 2
MG_API int mg_smooth_lexgs(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx,
                    double hy, double omega, int sweeps, int mode, int dtype, void* stream);

/* Coarsest-grid solve: up to max_iterations x [one lexicographic GS sweep (omega), residual
 * r = f - coefficient*lap_h(u), norm = sqrt(hx*hy*sum_all r^2)], stopping as soon as
 * norm < tolerance.  One launch, one thread block; `info` (device, 2 doubles, may be NULL)
 * receives {sweeps done, last norm}.
 * Replaces IterativeSolver.solve as used by MultigridSolver._solve_coarse,
 * solvers/base.py:234-290 + solvers/multigrid.py:355-370. */
MG_API int mg_coarse_solve_lexgs(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f,
                          double hx, double hy, double omega, double coefficient, double tolerance,
                          int max_iterations, double* info, int dtype, void* stream);

/* Fine (nxf, nyf) -> coarse ((nxf-1)/2+1, (nyf-1)/2+1).  Coarse boundary = injection; interior:
 * full weighting 1/16-1/8-1/4, half weighting 1/8-1/2, or injection; no h^2 scaling.
 * Replaces RestrictionOperator.apply, operators/transfer.py:53-148. */
MG_API int mg_restrict(const void* fine, void* coarse, int nxf, int nyf, int64_t ld_f, int64_t ld_c,
                int method, int dtype_in, int dtype_out, void* stream);

/* Coarse (nxc, nyc) -> fine (2(nxc-1)+1, 2(nyc-1)+1), bilinear or injection, INCLUDING the
 * reference's treatment of the last fine row/column (odd points stay 0).  add = 0: fine = P c;
 * add = 1: fine += P c  (the `u += fine_correction` of solvers/multigrid.py:329).
 * Replaces ProlongationOperator.apply, operators/transfer.py:189-267. */
MG_API int mg_prolong(const void* coarse, void* fine, int nxc, int nyc, int64_t ld_c, int64_t ld_f,
               int method, int add, int dtype_in, int dtype_out, void* stream);

/* Deterministic sum of squares over ALL (nx, ny) points (fixed two-stage tree, no atomics):
 * out[0] = sum x^2 (fp64 accumulation).  `workspace` = device buffer of at least
 * mg_sumsq_workspace_doubles() doubles.  The caller forms sqrt(hx*hy*out[0]) = Grid.l2_norm,
 * core/grid.py:174-187. */
MG_API int mg_sumsq(const void* x, int nx, int ny, int64_t ld, int dtype, double* workspace, double* out,
             void* stream);
MG_API int mg_sumsq_workspace_doubles(void);

/* dst = (dtype_dst) src, element-wise over (nx, ny)  (PrecisionManager.convert_array,
 * core/precision.py:106-134). */
MG_API int mg_cast(const void* src, void* dst, int nx, int ny, int64_t ld_src, int64_t ld_dst,
            int dtype_src, int dtype_dst, void* stream);

/* y += alpha * x over (nx, ny) with independent dtypes (fp64 u += fp32 correction is the
 * defect-correction update; gpu/cuda_kernels.py:915-929). */
MG_API int mg_axpy(double alpha, const void* x, void* y, int nx, int ny, int64_t ld_x, int64_t ld_y,
            int dtype_x, int dtype_y, void* stream);

/* x = 0 over the full pitched (nx, ld) extent. */
MG_API int mg_zero(void* x, int nx, int64_t ld, int dtype, void* stream);

/* f[i][j] = amplitude * sin(kx*pi*x_i) * sin(ky*pi*y_j), x_i = x0 + i*(x1-x0)/(nx-1) evaluated in
 * fp64 and rounded to dtype (synthetic manufactured-solution data generated in HBM; the
 * README problem README.md:77-78 is amplitude = 2*pi^2, kx = ky = 1). */
MG_API int mg_fill_sinsin(void* f, int nx, int ny, int64_t ld, double x0, double x1, double y0, double y1,
                   double amplitude, double kx, double ky, int dtype, void* stream);

/* out[0] = max_ij |u[i][j] - amplitude*sin(kx*pi*x_i)*sin(ky*pi*y_j)|  (MMS error without a host
 * copy of u).  workspace as for mg_sumsq. */
MG_API int mg_maxerr_sinsin(const void* u, int nx, int ny, int64_t ld, double x0, double x1, double y0,
                     double y1, double amplitude, double kx, double ky, int dtype, double* workspace,
                     double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
