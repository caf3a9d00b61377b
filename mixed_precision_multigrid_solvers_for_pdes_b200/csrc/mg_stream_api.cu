// libmgb200: C ABI of the fused / temporally blocked V-cycle passes (mg_vc_* family).
#include <cudaTypedefs.h>
#include <string.h>
#include "mg_stream_inst.cuh"
#include "mg_stream_dd.cuh"

namespace mg {
namespace stream {
int launch_pass_f32_tma(int, int, int, bool, int, const Maps&, PassParams&, const StencilScalars<float>&, cudaStream_t);
int launch_pass_f32_cpa(int, int, int, bool, int, const Maps&, PassParams&, const StencilScalars<float>&, cudaStream_t);
int launch_pass_f64_tma(int, int, int, bool, int, const Maps&, PassParams&, const StencilScalars<double>&, cudaStream_t);
int launch_pass_f64_cpa(int, int, int, bool, int, const Maps&, PassParams&, const StencilScalars<double>&, cudaStream_t);
int launch_defect_down(bool, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, DDParams&,
                       const StencilScalars<double>&, const StencilScalars<float>&, cudaStream_t);
int defect_down_box_rows();
int launch_pass_f32_var(int, int, int, bool, const Maps&, PassParams&, const StencilScalars<float>&, cudaStream_t);
int launch_pass_f64_var(int, int, int, bool, const Maps&, PassParams&, const StencilScalars<double>&, cudaStream_t);

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// (nx, ny) row-major field with pitch ld  ->  2-D tiled map, box = RB rows x 128 columns, zero OOB fill
static int make_map(CUtensorMap* m, const void* base, int nx, int ny, int64_t ld, int dtype, int box_w = STRIP,
                    int box_h = RB) {
  auto enc = get_encode();
  if (!enc) return MG_ERR_UNSUPPORTED;
  const size_t esz = dtype == MG_F64 ? 8 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)ny, (cuuint64_t)nx};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dtype == MG_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MG_OK : MG_ERR_BADARG;
}

// halo of a pass with `ns` pipeline stages (Geometry<NS, BACK>::H)
static inline int halo(int ns, int back) { return ((ns + (back ? 2 : 0)) + 1) & ~1; }
static inline int num_strips(int ny, int ns, int back) {
  const int h = halo(ns, back);
  const int lo = h < 4 ? 4 : h, hi = STRIP - 1 - lo, stride = hi - lo + 1;
  // strip k owns global columns up to stride*k - 4 + hi
  int k = 0;
  while ((int64_t)stride * k - 4 + hi < ny - 1) ++k;
  return k + 1;
}
}  // namespace stream
}  // namespace mg

using namespace mg;
using namespace mg::stream;

// Shared argument checking + launch for every mg_vc_* entry point.
static int run_pass(const void* u_in, void* u_out, const void* f, const void* coarse_in, void* coarse_out,
                    const float* fine_in, float* resid_out, double* sumsq_out, double* workspace, int nx, int ny,
                    int64_t ld_in, int64_t ld_out, int64_t ld_f, int64_t ld_ci, int64_t ld_co, int64_t ld_fi,
                    int64_t ld_ro, double hx, double hy, double omega, double coefficient, int sweeps, int dtype,
                    int front, int back, int flags, void* stream, const char* what, int norm_lo = 0,
                    int norm_hi = -1, double shift = 0.0, const void* a = nullptr, int64_t ld_a = 0) {
  const bool varcoef = a != nullptr;  // -div(a grad u) + shift*u with the nodal coefficient field a
  const bool cpa = (flags & MG_VC_LOADER_CPASYNC) != 0;
  const bool store = (flags & MG_VC_NO_STORE) == 0;
  const bool u_zero = (flags & MG_VC_U_ZERO) != 0;
  const int smooth = (flags & MG_VC_JACOBI) ? SMOOTH_JACOBI : SMOOTH_RBGS;
  const int ns = num_stages(smooth, sweeps);  // pipeline stages of the smoothing part
  const int rows_override = (flags >> 8) & 0xFFF;
  if (dtype != MG_F32 && dtype != MG_F64) return MG_ERR_DTYPE;
  if (!f || nx < 3 || ny < 3 || ld_f < ny || hx <= 0 || hy <= 0 || !(shift >= 0.0)) return MG_ERR_BADARG;
  if (!u_zero && (!u_in || ld_in < ny)) return MG_ERR_BADARG;
  if (sweeps < 0 || sweeps > 2) return MG_ERR_UNSUPPORTED;
  if (varcoef) {
    // red-black GS, TMA-staged; fp64 passes carry one sweep (register budget of the three row windows)
    if (cpa || smooth != SMOOTH_RBGS || (dtype == MG_F64 && sweeps > 1) || ld_a < ny) return MG_ERR_UNSUPPORTED;
    if (front == FRONT_PROLONG && back == BACK_RESTRICT) return MG_ERR_UNSUPPORTED;
  }
  if (smooth == SMOOTH_JACOBI && cpa && sweeps > 0) return MG_ERR_UNSUPPORTED;  // Jacobi passes are TMA-staged only
  if (store && (!u_out || ld_out < ny || u_out == u_in)) return MG_ERR_BADARG;
  if (!store && back == BACK_NONE) return MG_ERR_BADARG;
  if (sweeps == 0 && front == FRONT_NONE && back == BACK_NONE) return MG_ERR_BADARG;
  const int nxc = (nx - 1) / 2 + 1, nyc = (ny - 1) / 2 + 1;
  if (front == FRONT_PROLONG || back == BACK_RESTRICT) {
    // ny must be odd.  An even nx is accepted: it is a row slab of a larger grid whose last local row is a ghost
    // row (coarse local row ic <-> fine local row 2ic, ic <= (nx-1)/2).
    if ((ny - 1) % 2) return MG_ERR_BADARG;
    if (front == FRONT_PROLONG && (!coarse_in || ld_ci < nyc)) return MG_ERR_BADARG;
    if (back == BACK_RESTRICT && (!coarse_out || ld_co < nyc)) return MG_ERR_BADARG;
  }
  if (front == FRONT_ADDFINE && (!fine_in || ld_fi < ny || dtype != MG_F64)) return MG_ERR_BADARG;
  if (back == BACK_RESID && (!resid_out || ld_ro < ny || dtype != MG_F64)) return MG_ERR_BADARG;
  const bool norm = back == BACK_NORM || back == BACK_RESID;
  if (norm && (!sumsq_out || !workspace)) return MG_ERR_BADARG;
  // mg_vc_workspace_doubles sizes the partials for tiles of >= 8 rows: a smaller override would overrun it
  if (norm && rows_override > 0 && rows_override < 8) return MG_ERR_BADARG;
  const size_t esz = dtype == MG_F64 ? 8 : 4;
  auto misaligned = [&](const void* p, int64_t ld, size_t es) {
    return p && (((uintptr_t)p & 15u) || (ld % (int64_t)(16 / es)));
  };
  if ((!u_zero && misaligned(u_in, ld_in, esz)) || misaligned(f, ld_f, esz) || (store && misaligned(u_out, ld_out, esz)) ||
      (front == FRONT_PROLONG && misaligned(coarse_in, ld_ci, esz)) ||
      (back == BACK_RESTRICT && misaligned(coarse_out, ld_co, esz)) ||
      (front == FRONT_ADDFINE && misaligned(fine_in, ld_fi, 4)) || (back == BACK_RESID && misaligned(resid_out, ld_ro, 4)) ||
      (varcoef && misaligned(a, ld_a, esz)))
    return MG_ERR_ALIGN;

  PassParams p;
  memset(&p, 0, sizeof(p));
  p.u_in = u_zero ? f : u_in;  // never dereferenced when u_zero
  p.u_out = u_out; p.f = f; p.coarse_in = coarse_in; p.coarse_out = coarse_out;
  p.fine_in = fine_in; p.resid_out = resid_out;
  p.a = a; p.ld_a = ld_a;
  p.partials = workspace;
  p.nx = nx; p.ny = ny; p.nxc = nxc; p.nyc = nyc;
  p.ld_in = u_zero ? ld_f : ld_in; p.ld_out = ld_out; p.ld_f = ld_f; p.ld_ci = ld_ci; p.ld_co = ld_co;
  p.ld_fi = ld_fi; p.ld_ro = ld_ro;
  p.u_zero = u_zero ? 1 : 0;
  p.norm_row_lo = norm_lo < 0 ? 0 : norm_lo;
  p.norm_row_hi = (norm_hi < 0 || norm_hi > nx) ? nx : norm_hi;
  p.nstrips = num_strips(ny, ns, back);
  p.rows_per_tile = rows_override > 0 ? ((rows_override + 1) & ~1) : 0;  // 0: the launcher picks (pick_rows)
  p.tile_overlap = 2 * halo(ns, back) + 2;
  p.store_u = store ? 1 : 0;

  Maps m;
  memset(&m, 0, sizeof(m));
  if (!cpa) {
    int rc = MG_OK;
    if (!u_zero) rc = make_map(&m.u, u_in, nx, ny, ld_in, dtype);
    if (rc == MG_OK) rc = make_map(&m.f, f, nx, ny, ld_f, dtype);
    if (rc == MG_OK && front == FRONT_ADDFINE) rc = make_map(&m.e, fine_in, nx, ny, ld_fi, MG_F32);
    if (rc == MG_OK && front == FRONT_PROLONG)  // the coarse correction rides the TMA ring (see StageCoarse)
      rc = make_map(&m.e, coarse_in, nxc, nyc, ld_ci, dtype, COARSE_BOX_W, RB / 2 + 1);
    if (rc == MG_OK && varcoef) rc = make_map(&m.a, a, nx, ny, ld_a, dtype);
    if (rc != MG_OK) return rc;
  }
  cudaStream_t st = as_stream(stream);
  // isotropic, unrelaxed: the 5-instruction point update (bit-identical to the general one on dyadic grids without
  // a shift; both pin every rounding, so results never depend on the tiling or the slab decomposition)
  const bool simple = (omega == 1.0) && (hx == hy);
  int rc;
  if (varcoef) {
    // the variable-coefficient point update scales by 0.5/h^2 (the halving of the face means folded in) and takes
    // `shift` as is; `noblend` = (omega == 1)
    if (dtype == MG_F64) {
      auto sc = make_scalars<double>(hx, hy, omega, coefficient, shift);
      sc.ihx2 = 0.5 / sc.hx2; sc.ihy2 = 0.5 / sc.hy2;
      rc = launch_pass_f64_var(sweeps, front, back, omega == 1.0, m, p, sc, st);
    } else {
      auto sc = make_scalars<float>(hx, hy, omega, coefficient, shift);
      sc.ihx2 = 0.5f / sc.hx2; sc.ihy2 = 0.5f / sc.hy2;
      rc = launch_pass_f32_var(sweeps, front, back, omega == 1.0, m, p, sc, st);
    }
  } else if (dtype == MG_F64) {
    auto sc = make_scalars<double>(hx, hy, omega, coefficient, shift);
    rc = cpa ? launch_pass_f64_cpa(sweeps, front, back, simple, smooth, m, p, sc, st)
             : launch_pass_f64_tma(sweeps, front, back, simple, smooth, m, p, sc, st);
  } else {
    auto sc = make_scalars<float>(hx, hy, omega, coefficient, shift);
    rc = cpa ? launch_pass_f32_cpa(sweeps, front, back, simple, smooth, m, p, sc, st)
             : launch_pass_f32_tma(sweeps, front, back, simple, smooth, m, p, sc, st);
  }
  if (rc != MG_OK) return rc;
  if (norm) {
    const int ntiles = (nx + p.rows_per_tile - 1) / p.rows_per_tile;
    const int n = ((p.nstrips + WARPS - 1) / WARPS) * WARPS * ntiles;
    reduce_partials_sum(workspace, n, sumsq_out, st);
  }
  return check_launch(what, norm ? 2 : 1);
}

extern "C" {

int mg_vc_workspace_doubles(int nx, int ny) {
  if (nx < 3 || ny < 3) return 0;
  const int nstrips = num_strips(ny, 6, 1) + WARPS;  // widest halo (fused defect + down pass: 8) => most strips
  const int ntiles = (nx + 7) / 8;
  return nstrips * ntiles + 8;
}

int mg_vc_pass(const void* u_in, void* u_out, const void* f, const void* coarse_in, void* coarse_out,
               double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out, int64_t ld_f,
               int64_t ld_ci, int64_t ld_co, double hx, double hy, double omega, double coefficient, int sweeps,
               int dtype, int flags, void* stream) {
  if ((flags & MG_VC_RESTRICT) && (flags & MG_VC_NORM)) return MG_ERR_UNSUPPORTED;
  const int front = (flags & MG_VC_PROLONG) ? FRONT_PROLONG : FRONT_NONE;
  const int back = (flags & MG_VC_RESTRICT) ? BACK_RESTRICT : ((flags & MG_VC_NORM) ? BACK_NORM : BACK_NONE);
  return run_pass(u_in, u_out, f, coarse_in, coarse_out, nullptr, nullptr, sumsq_out, workspace, nx, ny, ld_in, ld_out,
                  ld_f, ld_ci, ld_co, 0, 0, hx, hy, omega, coefficient, sweeps, dtype, front, back, flags, stream,
                  "mg_vc_pass");
}

int mg_vc_pass_slab(const void* u_in, void* u_out, const void* f, const void* coarse_in, void* coarse_out,
                    double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out, int64_t ld_f,
                    int64_t ld_ci, int64_t ld_co, double hx, double hy, double omega, double coefficient, int sweeps,
                    int dtype, int flags, int norm_row_lo, int norm_row_hi, double shift, void* stream) {
  if ((flags & MG_VC_RESTRICT) && (flags & MG_VC_NORM)) return MG_ERR_UNSUPPORTED;
  const int front = (flags & MG_VC_PROLONG) ? FRONT_PROLONG : FRONT_NONE;
  const int back = (flags & MG_VC_RESTRICT) ? BACK_RESTRICT : ((flags & MG_VC_NORM) ? BACK_NORM : BACK_NONE);
  return run_pass(u_in, u_out, f, coarse_in, coarse_out, nullptr, nullptr, sumsq_out, workspace, nx, ny, ld_in, ld_out,
                  ld_f, ld_ci, ld_co, 0, 0, hx, hy, omega, coefficient, sweeps, dtype, front, back, flags, stream,
                  "mg_vc_pass_slab", norm_row_lo, norm_row_hi, shift);
}

int mg_vc_defect_pass_slab(const void* u_in, void* u_out, const void* f, const void* e_in, void* r_out,
                           double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out,
                           int64_t ld_f, int64_t ld_e, int64_t ld_r, double hx, double hy, double coefficient, int flags,
                           int norm_row_lo, int norm_row_hi, double shift, void* stream) {
  const int front = e_in ? FRONT_ADDFINE : FRONT_NONE;
  const int back = r_out ? BACK_RESID : BACK_NONE;
  // MG_VC_U_ZERO is honoured: the first defect passes of a solve from u = 0 neither read nor memset the iterate
  int fl = flags & ~(MG_VC_PROLONG | MG_VC_RESTRICT | MG_VC_NORM);
  if (!e_in) fl |= MG_VC_NO_STORE;
  return run_pass(u_in, e_in ? u_out : nullptr, f, nullptr, nullptr, (const float*)e_in, (float*)r_out, sumsq_out,
                  workspace, nx, ny, ld_in, ld_out, ld_f, 0, 0, ld_e, ld_r, hx, hy, 1.0, coefficient, 0, MG_F64, front,
                  back, fl, stream, "mg_vc_defect_pass_slab", norm_row_lo, norm_row_hi, shift);
}

int mg_vc_defect_pass(const void* u_in, void* u_out, const void* f, const void* e_in, void* r_out, double* sumsq_out,
                      double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out, int64_t ld_f, int64_t ld_e,
                      int64_t ld_r, double hx, double hy, double coefficient, int flags, void* stream) {
  const int front = e_in ? FRONT_ADDFINE : FRONT_NONE;
  const int back = r_out ? BACK_RESID : BACK_NONE;
  int fl = flags & ~(MG_VC_PROLONG | MG_VC_RESTRICT | MG_VC_NORM);
  if (!e_in) fl |= MG_VC_NO_STORE;  // nothing to add: u is unchanged, only the residual is produced
  return run_pass(u_in, e_in ? u_out : nullptr, f, nullptr, nullptr, (const float*)e_in, (float*)r_out, sumsq_out,
                  workspace, nx, ny, ld_in, ld_out, ld_f, 0, 0, ld_e, ld_r, hx, hy, 1.0, coefficient, 0, MG_F64, front,
                  back, fl, stream, "mg_vc_defect_pass");
}

int mg_vcv_pass_slab(const void* u_in, void* u_out, const void* f, const void* a, const void* coarse_in, void* coarse_out,
                     double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out, int64_t ld_f,
                     int64_t ld_a, int64_t ld_ci, int64_t ld_co, double hx, double hy, double omega, int sweeps, int dtype,
                     int flags, int norm_row_lo, int norm_row_hi, double shift, void* stream) {
  if (!a) return MG_ERR_BADARG;
  if ((flags & MG_VC_RESTRICT) && (flags & MG_VC_NORM)) return MG_ERR_UNSUPPORTED;
  if (flags & (MG_VC_JACOBI | MG_VC_LOADER_CPASYNC)) return MG_ERR_UNSUPPORTED;
  const int front = (flags & MG_VC_PROLONG) ? FRONT_PROLONG : FRONT_NONE;
  const int back = (flags & MG_VC_RESTRICT) ? BACK_RESTRICT : ((flags & MG_VC_NORM) ? BACK_NORM : BACK_NONE);
  return run_pass(u_in, u_out, f, coarse_in, coarse_out, nullptr, nullptr, sumsq_out, workspace, nx, ny, ld_in, ld_out,
                  ld_f, ld_ci, ld_co, 0, 0, hx, hy, omega, 1.0, sweeps, dtype, front, back, flags, stream,
                  "mg_vcv_pass_slab", norm_row_lo, norm_row_hi, shift, a, ld_a);
}

int mg_vcv_defect_pass_slab(const void* u_in, void* u_out, const void* f, const void* a, const void* e_in, void* r_out,
                            double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in, int64_t ld_out,
                            int64_t ld_f, int64_t ld_a, int64_t ld_e, int64_t ld_r, double hx, double hy, int flags,
                            int norm_row_lo, int norm_row_hi, double shift, void* stream) {
  if (!a) return MG_ERR_BADARG;
  if (flags & (MG_VC_JACOBI | MG_VC_LOADER_CPASYNC)) return MG_ERR_UNSUPPORTED;
  const int front = e_in ? FRONT_ADDFINE : FRONT_NONE;
  const int back = r_out ? BACK_RESID : BACK_NONE;
  int fl = flags & ~(MG_VC_PROLONG | MG_VC_RESTRICT | MG_VC_NORM);
  if (!e_in) fl |= MG_VC_NO_STORE;
  return run_pass(u_in, e_in ? u_out : nullptr, f, nullptr, nullptr, (const float*)e_in, (float*)r_out, sumsq_out,
                  workspace, nx, ny, ld_in, ld_out, ld_f, 0, 0, ld_e, ld_r, hx, hy, 1.0, 1.0, 0, MG_F64, front, back, fl,
                  stream, "mg_vcv_defect_pass_slab", norm_row_lo, norm_row_hi, shift, a, ld_a);
}

int mg_vc_defect_down_pass_slab(const void* u_in, void* u_out, const void* f, const void* e_in, void* r_out, void* e_out,
                                void* coarse_out, double* sumsq_out, double* workspace, int nx, int ny, int64_t ld_in,
                                int64_t ld_out, int64_t ld_f, int64_t ld_e, int64_t ld_r, int64_t ld_eo, int64_t ld_co,
                                double hx, double hy, double omega, double coefficient, int flags, int norm_row_lo,
                                int norm_row_hi, double shift, void* stream) {
  const bool u_zero = (flags & MG_VC_U_ZERO) != 0, has_e = e_in != nullptr;
  const int rows_override = (flags >> 8) & 0xFFF;
  if (flags & (MG_VC_JACOBI | MG_VC_LOADER_CPASYNC)) return MG_ERR_UNSUPPORTED;
  if (!f || !r_out || !e_out || !coarse_out || !sumsq_out || !workspace || nx < 3 || ny < 3 || hx <= 0 || hy <= 0 ||
      !(shift >= 0.0) || ld_f < ny || ld_r < ny || ld_eo < ny)
    return MG_ERR_BADARG;
  if ((ny - 1) % 2) return MG_ERR_BADARG;  // an even nx is a row slab whose last local row is a ghost row
  const int nxc = (nx - 1) / 2 + 1, nyc = (ny - 1) / 2 + 1;
  if (ld_co < nyc) return MG_ERR_BADARG;
  if (!u_zero && (!u_in || ld_in < ny)) return MG_ERR_BADARG;
  if (has_e && (!u_out || ld_out < ny || ld_e < ny || u_out == u_in)) return MG_ERR_BADARG;
  if (rows_override > 0 && rows_override < 8) return MG_ERR_BADARG;
  auto mis = [&](const void* q, int64_t ld, size_t es) { return q && (((uintptr_t)q & 15u) || (ld % (int64_t)(16 / es))); };
  if ((!u_zero && mis(u_in, ld_in, 8)) || mis(f, ld_f, 8) || (has_e && (mis(u_out, ld_out, 8) || mis(e_in, ld_e, 4))) ||
      mis(r_out, ld_r, 4) || mis(e_out, ld_eo, 4) || mis(coarse_out, ld_co, 4))
    return MG_ERR_ALIGN;
  DDParams p;
  memset(&p, 0, sizeof(p));
  p.u_in = (const double*)(u_zero ? f : u_in);
  p.u_out = (double*)u_out; p.f = (const double*)f; p.e_in = (const float*)e_in;
  p.r_out = (float*)r_out; p.e_out = (float*)e_out; p.coarse_out = (float*)coarse_out; p.partials = workspace;
  p.nx = nx; p.ny = ny; p.nxc = nxc; p.nyc = nyc;
  p.ld_u = u_zero ? ld_f : ld_in; p.ld_uo = ld_out; p.ld_f = ld_f; p.ld_e = ld_e; p.ld_r = ld_r; p.ld_eo = ld_eo; p.ld_co = ld_co;
  p.u_zero = u_zero ? 1 : 0; p.has_e = has_e ? 1 : 0;
  p.norm_row_lo = norm_row_lo < 0 ? 0 : norm_row_lo;
  p.norm_row_hi = (norm_row_hi < 0 || norm_row_hi > nx) ? nx : norm_row_hi;
  p.nstrips = num_strips(ny, 6, 1);  // halo 8: see DDGeometry
  p.rows_per_tile = rows_override > 0 ? ((rows_override + 1) & ~1) : 0;  // 0: the launcher picks (pick_rows)
  CUtensorMap mu, mf, me;
  memset(&mu, 0, sizeof(mu)); memset(&me, 0, sizeof(me));
  int rc = MG_OK;
  const int box_rows = defect_down_box_rows();
  if (!u_zero) rc = make_map(&mu, u_in, nx, ny, ld_in, MG_F64, STRIP, box_rows);
  if (rc == MG_OK) rc = make_map(&mf, f, nx, ny, ld_f, MG_F64, STRIP, box_rows);
  if (rc == MG_OK && has_e) rc = make_map(&me, e_in, nx, ny, ld_e, MG_F32, STRIP, box_rows);
  if (rc != MG_OK) return rc;
  cudaStream_t st = as_stream(stream);
  const bool simple = (omega == 1.0) && (hx == hy);
  auto sd = make_scalars<double>(hx, hy, omega, coefficient, shift);
  auto sf = make_scalars<float>(hx, hy, omega, coefficient, shift);
  rc = launch_defect_down(simple, mu, mf, me, p, sd, sf, st);
  if (rc != MG_OK) return rc;
  const int ntiles = (nx + p.rows_per_tile - 1) / p.rows_per_tile;
  reduce_partials_sum(workspace, ((p.nstrips + WARPS - 1) / WARPS) * WARPS * ntiles, sumsq_out, st);
  return check_launch("mg_vc_defect_down_pass_slab", 2);
}

int mg_vc_smooth(const void* u_in, void* u_out, const void* f, int nx, int ny, int64_t ld_in, int64_t ld_out,
                 int64_t ld_f, double hx, double hy, double omega, int sweeps, int dtype, int flags, void* stream) {
  return mg_vc_pass(u_in, u_out, f, nullptr, nullptr, nullptr, nullptr, nx, ny, ld_in, ld_out, ld_f, 0, 0, hx, hy,
                    omega, 1.0, sweeps, dtype, flags & ~(MG_VC_PROLONG | MG_VC_RESTRICT | MG_VC_NORM | MG_VC_NO_STORE),
                    stream);
}

int mg_vc_residual_restrict(const void* u, const void* f, void* coarse_out, int nx, int ny, int64_t ld_u,
                            int64_t ld_f, int64_t ld_co, double hx, double hy, double coefficient, int dtype,
                            int flags, void* stream) {
  const int fl = (flags & ~(MG_VC_PROLONG | MG_VC_NORM)) | MG_VC_RESTRICT | MG_VC_NO_STORE;
  return mg_vc_pass(u, nullptr, f, nullptr, coarse_out, nullptr, nullptr, nx, ny, ld_u, 0, ld_f, 0, ld_co, hx, hy, 1.0,
                    coefficient, 0, dtype, fl, stream);
}

int mg_vc_prolong_correct_smooth(const void* u_in, void* u_out, const void* f, const void* coarse_in, int nx, int ny,
                                 int64_t ld_in, int64_t ld_out, int64_t ld_f, int64_t ld_ci, double hx, double hy,
                                 double omega, int sweeps, int dtype, int flags, void* stream) {
  const int fl = (flags & ~(MG_VC_RESTRICT | MG_VC_NORM | MG_VC_NO_STORE)) | MG_VC_PROLONG;
  return mg_vc_pass(u_in, u_out, f, coarse_in, nullptr, nullptr, nullptr, nx, ny, ld_in, ld_out, ld_f, ld_ci, 0, hx,
                    hy, omega, 1.0, sweeps, dtype, fl, stream);
}

}  // extern "C"
