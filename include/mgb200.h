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
/* process-wide count of CUDA kernels this library has launched (bench.py's gpu_launches) */
MG_API long long mg_launch_count(void);

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

/* In-place lexicographic Gauss-Seidel (i-major, then j): identical arithmetic to the sequential sweep at any size.
 * mode 0: forward sweeps  -- GaussSeidelSmoother(red_black=False).smooth, solvers/smoothers.py:153-173;
 * mode 1: backward sweeps (i, j descending)  -- _backward_sweep, solvers/smoothers.py:268-284;
 * mode 2: each sweep = forward then backward  -- SymmetricGaussSeidelSmoother.smooth, :246-266.
 * One launch per sweep direction, a skewed wavefront pipelined over warps: warp w owns 32 consecutive rows, lane l
 * relaxes column t - l at step t (new value of the row above by shuffle), and the first row of warp w follows the last
 * row of warp w-1 through a progress counter in global memory (dependencies only on lower block indices: no
 * co-residency requirement; a 2 s watchdog traps instead of hanging).  A sweep costs about ny + 2*nx dependent
 * steps of ~0.17 us (measured: 0.6 ms at 1025^2, 2.8 ms at 4097^2, 12 ms at 16385^2, profiles/r02_lexgs_bench.log)
 * instead of nx + ny block barriers of a one-block wavefront (which mg_coarse_solve_lexgs keeps for the coarsest
 * grid): the reference's setup default made usable at every size, not an HBM-bound kernel -- a red-black sweep of
 * 16385^2 takes 0.25 ms. */
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

/* Host read-back of n (<= 1024) device doubles WITHOUT the copy engine: dst_host is page-locked host memory
 * (cudaHostAlloc / torch pin_memory); a one-warp kernel stores the values through its device alias, so the read
 * of a residual norm (the one scalar a cycle returns to the host, solvers/base.py:123-143) never queues behind a
 * bulk device-to-host transfer running on another stream.  Asynchronous: synchronise `stream` before reading.
 * Falls back to cudaMemcpyAsync when dst_host is not mapped page-locked memory. */
MG_API int mg_read_doubles(const double* src, double* dst_host, int n, void* stream);

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

/* Zero the boundary ring of a field: first and last column always, first / last row when the flag is set (a row slab
 * holds them only on the physical boundary).  The reference's residual equals f on the ring and its norm sums over ALL
 * points (operators/laplacian.py:64,117; core/grid.py:187): a source that does not vanish there can never meet the
 * tolerance although those values enter no equation, so the facade clears the ring of f. */
MG_API int mg_zero_ring(void* x, int nx, int ny, int64_t ld, int first_row, int last_row, int dtype, void* stream);

/* Right-hand side of one theta-method heat step in ONE pass over HBM (fp64): the linear system the reference documents,
 *   (I - theta*dt*L_h) u^{n+1} = u^n + (1-theta)*dt*L_h u^n + dt*(theta f^{n+1} + (1-theta) f^n)
 * (docs/methodology.md:710; L_h = alpha*lap_h, or div(a grad .) with a nodal field a), divided by theta*dt:
 *   rhs = lam * (u + c_lap * L_h u + c_f1 * f1 + c_f0 * f0),
 * L_h u taken as 0 on the first / last local row and column, the boundary ring zeroed (first / last row only when the
 * flag says the slab touches the physical boundary), and sumsq_out[0] = sum of rhs^2 over rows [norm_row_lo,
 * norm_row_hi) for the step's relative stopping test (`workspace`: mg_sumsq_workspace_doubles() doubles).  f1, f0, a
 * may be NULL.  Replaces the eager copy / add / scale / norm sequence of the reference's time loop
 * (applications/heat_solver.py:308-390) and of round 1's driver. */
MG_API int mg_heat_rhs(const double* u, const double* f1, const double* f0, const double* a, double* rhs,
                double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_u, int64_t ld_f1, int64_t ld_f0,
                int64_t ld_a, int64_t ld_rhs, double hx, double hy, double c_lap, double c_f1, double c_f0, double lam,
                int zero_first_row, int zero_last_row, int norm_row_lo, int norm_row_hi, void* stream);

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

/* Helmholtz-shifted variants: operator coefficient*lap_h + shift (shift >= 0), i.e. (-lap + lambda) u = f for
 * coefficient = -1: the implicit heat step (I - alpha*dt*lap) u = rhs of docs/methodology.md:710 divided by
 * alpha*dt (the shifted stencil exists only in applications/heat_equation.py:459-497 of the reference, which
 * solves it with plain GS).  The smoother relaxes (-lap_h + shift) u = f.  shift = 0 reproduces the functions above. */
MG_API int mg_residual_h(const void* u, const void* f, void* r, int nx, int ny, int64_t ld_u, int64_t ld_f,
                  int64_t ld_r, double hx, double hy, double coefficient, double shift, int dtype_in,
                  int dtype_out, void* stream);
MG_API int mg_smooth_rbgs_h(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx,
                     double hy, double omega, double shift, int sweeps, int dtype, void* stream);
MG_API int mg_coarse_solve_lexgs_h(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f,
                            double hx, double hy, double omega, double coefficient, double shift,
                            double tolerance, int max_iterations, double* info, int dtype, void* stream);

/* Variable-coefficient operator A u = -div(a grad u) + shift*u with a nodal coefficient field `a` (same shape and
 * dtype as u; arithmetic-mean face coefficients).  The reference advertises this problem class (README.md:175) but
 * ships no operator for it: there is no reference arithmetic to mirror; a == 1, shift == 0 reduces to
 * coefficient = -1 of the functions above.
 * mg_varcoef_residual: r = f - A u (r = f on the boundary), or r = A u when apply_only != 0 (f may then be NULL).
 * mg_varcoef_smooth_rbgs: in-place red-black Gauss-Seidel relaxing A u = f. */
MG_API int mg_varcoef_residual(const void* u, const void* f, const void* a, void* r, int nx, int ny, int64_t ld_u,
                        int64_t ld_f, int64_t ld_a, int64_t ld_r, double hx, double hy, double shift,
                        int apply_only, int dtype, void* stream);
MG_API int mg_varcoef_smooth_rbgs(void* u, const void* f, const void* a, int nx, int ny, int64_t ld_u,
                           int64_t ld_f, int64_t ld_a, double hx, double hy, double shift, double omega,
                           int sweeps, int dtype, void* stream);

/* Coarsest-level solve for the variable-coefficient operator: <= max_iterations x [one red-black GS sweep, residual,
 * h-scaled L2 norm], stop below `tolerance` -- IterativeSolver.solve (solvers/base.py:258-285) with the red-black
 * variable-coefficient smoother as the solver, in ONE launch (stopping test in the kernel).  Grids up to 65 x 65.
 * `info` (device, 2 doubles, may be NULL) = {sweeps, last norm}. */
MG_API int mg_varcoef_coarse_solve(void* u, const void* f, const void* a, int nx, int ny, int64_t ld_u, int64_t ld_f,
                            int64_t ld_a, double hx, double hy, double shift, double omega, double tolerance,
                            int max_iterations, double* info, int dtype, void* stream);

/* The defect pass and the pre-smoothing pass of the error equation in ONE pass over HBM (round 2): mg_vc_defect_pass_slab
 * followed by mg_vc_pass_slab(MG_VC_U_ZERO | MG_VC_RESTRICT, sweeps = 2) -- the refinement loop of
 * docs/methodology.md:337-360 around the down leg of solvers/multigrid.py:286-300:
 *     u_out = u_in + (double) e_in                 (e_in NULL: u unchanged and not stored; MG_VC_U_ZERO: u_in == 0, not read)
 *     r_out = (float)(f - A u_out),  sumsq_out[0] = sum of the squared fp64 residual over rows [norm_row_lo, norm_row_hi)
 *     e_out = 2 red-black GS sweeps from zero on A e = r_out  (fp32; omega, coefficient, shift as in mg_vc_pass_slab)
 *     coarse_out = R_fw(r_out - A e_out)
 * The fp32 residual row is stored (the up pass needs it as its right-hand side) but consumed by the smoothing
 * pipeline from registers: 37 instead of 41 bytes per point and one launch less.  u_out, r_out, e_out and coarse_out
 * are bit-identical to the two separate passes.  TMA loader, red-black GS; fields 16-byte aligned. */
MG_API int mg_vc_defect_down_pass_slab(const void* u_in, void* u_out, const void* f, const void* e_in, void* r_out,
                                void* e_out, void* coarse_out, double* sumsq_out, double* workspace, int nx, int ny,
                                int64_t ld_in, int64_t ld_out, int64_t ld_f, int64_t ld_e, int64_t ld_r, int64_t ld_eo,
                                int64_t ld_co, double hx, double hy, double omega, double coefficient, int flags,
                                int norm_row_lo, int norm_row_hi, double shift, void* stream);

/* Fused / temporally blocked passes of the variable-coefficient operator  A u = -div(a grad u) + shift*u  (README.md:175
 * advertises the problem class, docs/methodology.md:710 the implicit heat system it serves; the reference ships no operator:
 * SURVEY 8f-1, parity unpinned).  Same pass structure, flags and slab semantics as mg_vc_pass_slab /
 * mg_vc_defect_pass_slab; `a` = nodal coefficient field of the level (same shape and dtype as u: +1 word per point of
 * traffic), staged through the same TMA ring.  Red-black GS only, TMA loader only, fp64 passes carry at most one sweep;
 * MG_VC_PROLONG together with MG_VC_RESTRICT is not instantiated.  On grids with power-of-two spacings the results are
 * bit-identical to mg_varcoef_smooth_rbgs / mg_varcoef_residual (true division by the per-point diagonal is kept). */
MG_API int mg_vcv_pass_slab(const void* u_in, void* u_out, const void* f, const void* a, const void* coarse_in,
                     void* coarse_out, double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in,
                     int64_t ld_out, int64_t ld_f, int64_t ld_a, int64_t ld_ci, int64_t ld_co, double hx, double hy,
                     double omega, int sweeps, int dtype, int flags, int norm_row_lo, int norm_row_hi, double shift,
                     void* stream);
MG_API int mg_vcv_defect_pass_slab(const void* u_in, void* u_out, const void* f, const void* a, const void* e_in,
                            void* r_out, double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in,
                            int64_t ld_out, int64_t ld_f, int64_t ld_a, int64_t ld_e, int64_t ld_r, double hx, double hy,
                            int flags, int norm_row_lo, int norm_row_hi, double shift, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused, temporally blocked V/W-cycle passes ("vector path": TMA-staged, 16-byte aligned fields)
 *
 * One launch performs, in one pass over HBM and out of place (u_in -> u_out, u_out != u_in):
 *     [MG_VC_PROLONG: u += bilinear P(coarse_in)]                 (transfer.py:234-267, multigrid.py:329)
 *     -> `sweeps` (0..2) red-black Gauss-Seidel sweeps            (smoothers.py:175-207)
 *        or, with MG_VC_JACOBI, damped-Jacobi sweeps              (smoothers.py:41-86)
 *     -> [MG_VC_RESTRICT: coarse_out = full-weighting R(f - A u)] (laplacian.py:105-124, transfer.py:100-124)
 *        or [MG_VC_NORM: sumsq_out[0] = sum over all points of (f - A u)^2]  (grid.py:174-187)
 * so a V(2,2) level costs three passes: smooth+residual+restrict going down, prolong+correct+smooth
 * (+norm on the finest level) going up  --  the kernels named in BASELINE.json's north star.
 * Arithmetic uses reciprocal multiplies and FMA: bit-identical to the basic kernels when hx^2, hy^2
 * are powers of two (n = 2^k+1 on the unit square), within a few ulp otherwise.
 * ------------------------------------------------------------------------------------------- */
#define MG_VC_PROLONG 1          /* front stage: add the prolongated coarse correction */
#define MG_VC_RESTRICT 2         /* back stage: residual + full-weighting restriction into coarse_out */
#define MG_VC_NORM 4             /* back stage: residual sum of squares into sumsq_out (needs workspace) */
#define MG_VC_LOADER_CPASYNC 16  /* stage rows with per-lane cp.async instead of TMA */
#define MG_VC_NO_STORE 32        /* do not write u_out (pure residual passes, sweeps = 0) */
#define MG_VC_U_ZERO 64          /* u_in is identically zero and is not read (u_in may be NULL): the first
                                    pre-smoothing pass of every coarse-level / error-equation solve */
#define MG_VC_JACOBI 128         /* the `sweeps` are damped-Jacobi sweeps (JacobiSmoother / WeightedJacobiSmoother,
                                    smoothers.py:41-86, 210-225; omega = relaxation parameter) instead of red-black
                                    GS: one pipeline stage per sweep, previous-iterate rows kept in registers.
                                    TMA loader only (MG_ERR_UNSUPPORTED together with MG_VC_LOADER_CPASYNC) */
#define MG_VC_ROWS(r) (((r) & 0xFFF) << 8) /* override rows per tile (0 = auto) */

/* doubles of workspace MG_VC_NORM needs for an (nx, ny) field */
MG_API int mg_vc_workspace_doubles(int nx, int ny);

MG_API int mg_vc_pass(const void* u_in, void* u_out, const void* f, const void* coarse_in, void* coarse_out,
               double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out,
               int64_t ld_f, int64_t ld_ci, int64_t ld_co, double hx, double hy, double omega,
               double coefficient, int sweeps, int dtype, int flags, void* stream);

/* Mixed-precision defect-correction pass on the fp64 iterate ("fp32 smoothing / fp64 residual"; the
 * reference's mixed_precision_residual_kernel + mixed_precision_correction_kernel,
 * gpu/cuda_kernels.py:843-929, and the refinement loop of docs/methodology.md:337-360), one HBM pass:
 *     u_out = u_in + (double) e_in            (e_in: fp32 fine-grid correction; NULL: u unchanged, no store)
 *     r_out = (float) (f - coefficient*lap_h u_out)   (r_out: fp32; NULL: no residual stage)
 *     sumsq_out[0] = sum over all points of the fp64 residual squared (needs workspace)
 * u, f are fp64; out of place (u_out != u_in).  flags: MG_VC_LOADER_CPASYNC, MG_VC_ROWS, and MG_VC_U_ZERO (u_in is
 * identically zero and is not read: the first passes of a solve from the zero initial guess, multigrid.py:208-211,
 * then need neither a memset of the iterate nor its 8 bytes per point of read traffic). */
MG_API int mg_vc_defect_pass(const void* u_in, void* u_out, const void* f, const void* e_in, void* r_out,
                      double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out,
                      int64_t ld_f, int64_t ld_e, int64_t ld_r, double hx, double hy, double coefficient,
                      int flags, void* stream);

/* Row-slab variants for the 1-D domain decomposition over GPUs (the reference's strip decomposition,
 * gpu/multi_gpu_solver.py:347-383): the field handed in is a slab of rows of a larger grid including its
 * ghost rows; the first / last local rows are treated like Dirichlet rows (they are ghost rows refreshed by
 * the halo exchange, or true boundary rows on the first / last rank).  nx may be even.  Only rows
 * [norm_row_lo, norm_row_hi) enter the residual sum (each rank sums the rows it owns; hi < 0: all).
 * `shift` >= 0 adds the Helmholtz term (operator coefficient*lap_h + shift; 0 = Poisson). */
MG_API int mg_vc_pass_slab(const void* u_in, void* u_out, const void* f, const void* coarse_in, void* coarse_out,
                    double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out,
                    int64_t ld_f, int64_t ld_ci, int64_t ld_co, double hx, double hy, double omega,
                    double coefficient, int sweeps, int dtype, int flags, int norm_row_lo, int norm_row_hi,
                    double shift, void* stream);
MG_API int mg_vc_defect_pass_slab(const void* u_in, void* u_out, const void* f, const void* e_in, void* r_out,
                           double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in,
                           int64_t ld_out, int64_t ld_f, int64_t ld_e, int64_t ld_r, double hx, double hy,
                           double coefficient, int flags, int norm_row_lo, int norm_row_hi, double shift,
                           void* stream);

/* `sweeps` temporally blocked RB-GS sweeps in one HBM pass (replaces GaussSeidelSmoother(red_black=True)
 * .smooth, smoothers.py:117-151; SmoothingKernels.red_black_gauss_seidel / block_gauss_seidel_kernel,
 * gpu/cuda_kernels.py:348-390, 982-1048). */
MG_API int mg_vc_smooth(const void* u_in, void* u_out, const void* f, int nx, int ny, int64_t ld_in,
                 int64_t ld_out, int64_t ld_f, double hx, double hy, double omega, int sweeps, int dtype,
                 int flags, void* stream);

/* coarse_out = R_fw(f - coefficient*lap_h u) without materialising the fine residual (replaces
 * LaplacianOperator.residual + RestrictionOperator.apply, multigrid.py:294-300; TransferKernels.
 * compute_residual + .restriction, gpu/cuda_kernels.py:738-828). */
MG_API int mg_vc_residual_restrict(const void* u, const void* f, void* coarse_out, int nx, int ny, int64_t ld_u,
                            int64_t ld_f, int64_t ld_co, double hx, double hy, double coefficient,
                            int dtype, int flags, void* stream);

/* u_out = smooth^sweeps(u_in + P coarse_in) (replaces ProlongationOperator.apply + `u += e` + post-smooth,
 * multigrid.py:321-335; TransferKernels.prolongation, gpu/cuda_kernels.py:766-792). */
MG_API int mg_vc_prolong_correct_smooth(const void* u_in, void* u_out, const void* f, const void* coarse_in,
                                 int nx, int ny, int64_t ld_in, int64_t ld_out, int64_t ld_f,
                                 int64_t ld_ci, double hx, double hy, double omega, int sweeps, int dtype,
                                 int flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The coarse end of a cycle in one launch: the complete V (cycle = 0), W (1) or F (2) sub-cycle of
 * solvers/multigrid.py:253-337 over `nlev` levels starting at the (nx, ny) entry level, in place on u, by ONE
 * thread block holding every level in shared memory (levels: red-black GS `pre`/`post` sweeps, residual + full
 * weighting, bilinear prolongation + correction; coarsest level: lexicographic-GS solve to `coarse_tolerance`,
 * solvers/base.py:258-285).  Levels 0..nlev-2 are `dtype`, the coarsest `coarse_dtype` (fp32 levels with an fp64
 * coarsest level mirror the reference, which never converts the coarsest level, multigrid.py:270-272).
 * Usable when mg_small_cycle_smem_bytes(...) <= 200 KiB (e.g. 129^2 fp32 or 65^2 fp64 entry levels);
 * returns MG_ERR_UNSUPPORTED otherwise.  A W-cycle at 32769^2 visits the coarsest grid 8192 times: this
 * kernel replaces ~12 launches per visit by one.  `info` (device, 2 doubles, may be NULL) = {sweeps, norm} of
 * the last coarse solve; u_zero bit 0: start from u = 0 without reading u; bit 1 (profiling aid): `info` holds 16
 * doubles and the kernel ADDS SM clock cycles to info[2..5] = {smoothing, residual + restriction, prolongation, coarsest
 * solve} of the whole-block levels, info[6..9] = the same for the levels worked on by warp 0 alone (grids up to
 * 17 x 17), info[10..12] = {load, cycle, store}, info[13] = launches. */
MG_API int mg_small_cycle_smem_bytes(int nx, int ny, int nlev, int dtype, int coarse_dtype);
MG_API int mg_small_cycle(void* u, const void* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double hx, double hy,
                   int nlev, int cycle, int pre, int post, double omega, double coefficient, double shift,
                   double coarse_tolerance, int coarse_max_iterations, int u_zero, double* info, int dtype,
                   int coarse_dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The reference's SECONDARY solver, CorrectedMultigridSolver (solvers/corrected_multigrid.py:24-418; the V-cycle of
 * validation/simple_validation.py, mms_validation.py and the tutorials).  fp64, square spacing h, homogeneous
 * Dirichlet ring; every kernel keeps the reference's operand order without fused multiply-add, so a solve reproduces
 * the reference's solution bit for bit (tests/test_gpu_corrected.py against tests/golden/corrected_golden.npz).
 *   mg_cm_gs           `sweeps` lexicographic GS sweeps u[i,j] = 0.25*(u[i-1,j] + u[i+1,j] + u[i,j-1] + u[i,j+1]
 *                      + h^2 f[i,j]) in place (_gauss_seidel_iteration, :248-277), pipelined over warps like
 *                      mg_smooth_lexgs; the ring is not touched
 *   mg_cm_residual     r = f - (-lap_h u) inside, 0 on the ring (_compute_residual, :279-308); r may be NULL;
 *                      sumsq_out[0] = sum r^2 = the square of _compute_residual_norm (:310-316) when given together
 *                      with `workspace` (device, mg_cm_workspace_doubles() doubles)
 *   mg_cm_diff_sumsq   sumsq_out[0] = sum (a - b)^2  (the ||u - u_old|| test of _solve_coarsest, :384-387)
 *   mg_cm_restrict     full weighting / 16 on interior coarse points, 0 on the coarse ring (_restrict, :318-335)
 *   mg_cm_prolong_add  u += P(coarse) with the textbook bilinear prolongation (_prolongate, :337-364), then the ring of
 *                      u zeroed (_apply_boundary_conditions, :392-397) */
MG_API int mg_cm_workspace_doubles(void);
MG_API int mg_cm_gs(double* u, const double* f, int nx, int ny, int64_t ld_u, int64_t ld_f, double h, int sweeps,
                    void* stream);
MG_API int mg_cm_residual(const double* u, const double* f, double* r, double* sumsq_out, double* workspace, int nx,
                          int ny, int64_t ld_u, int64_t ld_f, int64_t ld_r, double h, void* stream);
MG_API int mg_cm_diff_sumsq(const double* a, const double* b, double* sumsq_out, double* workspace, int nx, int ny,
                            int64_t ld_a, int64_t ld_b, void* stream);
MG_API int mg_cm_restrict(const double* fine, double* coarse, int nxf, int nyf, int nxc, int nyc, int64_t ld_f,
                          int64_t ld_c, void* stream);
MG_API int mg_cm_prolong_add(const double* coarse, double* u, int nxc, int nyc, int nxf, int nyf, int64_t ld_c,
                             int64_t ld_u, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGB200_H */
