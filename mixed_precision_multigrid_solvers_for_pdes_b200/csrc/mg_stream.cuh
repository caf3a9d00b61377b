// libmgb200: the temporally blocked, fused red-black Gauss-Seidel streaming kernel (sm_100a).
//
// One kernel template covers the three fused passes of a V/W-cycle level:
//
//   [FRONT: u += P e_c]  ->  NU x (red half-sweep, black half-sweep)  ->  [BACK: r = f - A u, then
//                            or NU damped-Jacobi sweeps            full-weighting restriction  OR  sum r^2]
//
// in ONE pass over HBM: u_in and f are read once, u_out written once (plus 1/4-size coarse
// traffic).  Out of place (u_in -> u_out), so tiles never race on halos.
//
// Decomposition.  Every WARP owns a strip of 128 columns (32 lanes x 4 contiguous elements) and
// streams down a tile of rows.  The lane keeps a sliding window of rows in REGISTERS; stage s of
// the 2*NU half-sweeps works on the row loaded s steps ago, so when row i arrives, row i-2NU is
// final.  North/south neighbours are the lane's own registers, west/east neighbours of the lane's
// first/last element come from the adjacent lane by one warp shuffle per half-sweep.  Nothing is
// exchanged between warps: strips overlap by 2H columns (H = 2NU [+2 with a BACK stage]) and the
// overlap is recomputed (the recomputed loads hit in L2, DRAM traffic stays algorithmic).  Row
// tiles overlap the same way (H rows of warm-up).
//
// Data movement.  Rows are staged global -> shared by TMA (cp.async.bulk.tensor.2d, one elected
// lane per warp, mbarrier completion) in boxes of RB rows x 128 columns through an NSTAGE ring
// that is PRIVATE to the warp (no block-level barrier in the main loop).  TMA zero-fills
// everything outside the (nx, ny) extent, which gives the Dirichlet ring / padding handling for
// free.  LOADER = 1 swaps TMA for per-lane 16-byte cp.async with zero-fill (same ring).
// Results leave through 16-byte vector stores straight from registers.
//
// Arithmetic.  Multiplications by the precomputed 1/hx^2, 1/hy^2, 1/(2/hx^2+2/hy^2) and FMA
// contraction replace the reference's divisions.  On grids whose spacings are powers of two
// (n = 2^k+1 on the unit square: every BASELINE config) all those products are exact, so the
// results are BIT-IDENTICAL to the strict kernels / the reference; otherwise they differ by a few
// ulp (bar: 1e-12 relative per application).
#pragma once
#include <cuda.h>
#include "mg_common.cuh"

namespace mg {
namespace stream {

constexpr int LANE_V = 4;     // elements per lane per row
constexpr int STRIP = 128;    // columns per warp strip
constexpr int BACK_NONE = 0, BACK_RESTRICT = 1, BACK_NORM = 2, BACK_RESID = 3;  // RESID: store fp32 residual + norm
constexpr int FRONT_NONE = 0, FRONT_PROLONG = 1, FRONT_ADDFINE = 2;             // ADDFINE: u += (T)e, e fp32, fine grid
constexpr int LOADER_TMA = 0, LOADER_CPASYNC = 1;
// fp32 prolongation passes stage the coarse correction through the TMA ring too: RB/2 + 1 coarse rows of
// COARSE_BOX_W columns per box of RB fine rows: 65 are needed, the box start is rounded down to a multiple of 4
// columns (a TMA box must start 16-byte aligned in the inner dimension: an 8-byte aligned start raised "illegal
// instruction" on sm_100a), so up to 67 + 1 = 68 columns = 272 bytes per row
constexpr int COARSE_BOX_W = 68;
// staged for fp32 always; for fp64 only when a BACK stage already limits the kernel to 2 blocks per SM by registers
template <typename T, int FRONT, int BACK, int LOADER> struct StageCoarse {
  static constexpr bool value = (FRONT == FRONT_PROLONG) && (LOADER == LOADER_TMA) && (sizeof(T) == 4 || BACK != BACK_NONE);
};

constexpr int SMOOTH_RBGS = 0, SMOOTH_JACOBI = 1;
// pipeline stages of `nu` sweeps: a red-black sweep is two half-sweep stages, a Jacobi sweep is one
constexpr int num_stages(int smooth, int nu) { return smooth == SMOOTH_JACOBI ? nu : 2 * nu; }

// NS = number of pipeline stages; every stage widens the dependency cone by one row / column
template <int NS, int BACK> struct Geometry {
  static constexpr int H = ((NS + (BACK != BACK_NONE ? 2 : 0)) + 1) & ~1;  // halo, rounded up to even
  static constexpr int OWN_LO = H < 4 ? 4 : H;                    // local column 4 is global column g0+4
  static constexpr int OWN_HI = STRIP - 1 - OWN_LO;
  static constexpr int STRIDE = OWN_HI - OWN_LO + 1;              // multiple of 4
  static constexpr int ROW_LEAD = H;  // rows streamed before the first owned row (even: row parity, coarse mapping)
  static constexpr int ROW_TAIL = H;  // rows streamed after the last owned row
  static constexpr int WR = NS + 2 + (BACK != BACK_NONE ? 1 : 0);    // u window (ages 0..WR-1)
  static constexpr int FR = NS + 1 + (BACK != BACK_NONE ? 1 : 0);    // f window (ages 0..FR-1)
};

struct PassParams {
  const void* u_in;
  void* u_out;
  const void* f;
  const void* coarse_in;   // e_c (FRONT_PROLONG) or null
  void* coarse_out;        // restricted residual (BACK_RESTRICT) or null
  const void* a;           // nodal diffusion coefficient (VARCOEF kernels; cp.async loader reads it directly) or null
  const float* fine_in;    // fp32 fine-grid correction (FRONT_ADDFINE) or null
  float* resid_out;        // fp32 fine-grid residual (BACK_RESID) or null
  double* partials;        // one double per warp (BACK_NORM / BACK_RESID) or null
  int nx, ny, nxc, nyc;
  int64_t ld_in, ld_out, ld_f, ld_ci, ld_co, ld_fi, ld_ro, ld_a;
  int u_zero;              // 1: u_in is identically zero and is not read
  int norm_row_lo, norm_row_hi;  // rows [lo, hi) entering the residual sum (a slab sums only the rows it owns)
  int rows_per_tile;       // R (even); 0 on entry to the launcher = pick from the occupancy (pick_rows)
  int tile_overlap;        // lead + tail rows a tile streams besides its own
  int nstrips;
  int store_u;             // 0: do not write u_out (pure residual passes)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// 4 contiguous elements  <->  shared / global memory
template <typename T> struct Row4 { T v[4]; };
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

// Shared-memory reads take 32-bit shared-window addresses: through generic pointers the compiler rebuilt the window
// base (S2UR SR_CgaCtaId / ULEA chains) at every group of loads, ~14 % of the stall samples of the fp32 down pass (ncu r01).
__device__ __forceinline__ void lds4(uint32_t a, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
}
__device__ __forceinline__ void lds4(uint32_t a, double (&v)[4]) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(a));
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(a + 16u));
}
__device__ __forceinline__ void lds2(uint32_t a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a));
}
__device__ __forceinline__ void lds2(uint32_t a, double& x, double& y) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
__device__ __forceinline__ void lds1(uint32_t a, float& x) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); }
__device__ __forceinline__ void lds1(uint32_t a, double& x) { asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(a)); }

// Global stores predicated per lane, never branched: which lanes own which of their 4 columns differs along the warp
// (strip halos), and `if (own == ...) store` chains compiled to BSSY / jump table / BRX / BSYNC around every row
// (branch_resolving was the second largest stall of the fp32 down pass, ncu r01).
__device__ __forceinline__ void stg4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void stg4(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void stg4_if(bool on, float* p, float a, float b, float c, float d) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v4.f32 [%1], {%2, %3, %4, %5};\n\t}" ::"r"((int)on),
               "l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void stg4_if(bool on, double* p, double a, double b, double c, double d) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v2.f64 [%1], {%2, %3};\n\t@q st.global.v2.f64 [%1+16], {%4, %5};\n\t}" ::"r"(
          (int)on),
      "l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
      : "memory");
}
__device__ __forceinline__ void stg2_if(bool on, float* p, float a, float b) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v2.f32 [%1], {%2, %3};\n\t}" ::"r"((int)on), "l"(p), "f"(a),
               "f"(b)
               : "memory");
}
__device__ __forceinline__ void stg2_if(bool on, double* p, double a, double b) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.v2.f64 [%1], {%2, %3};\n\t}" ::"r"((int)on), "l"(p), "d"(a),
               "d"(b)
               : "memory");
}
__device__ __forceinline__ void stg1_if(bool on, float* p, float a) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.f32 [%1], %2;\n\t}" ::"r"((int)on), "l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void stg1_if(bool on, double* p, double a) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q st.global.f64 [%1], %2;\n\t}" ::"r"((int)on), "l"(p), "d"(a) : "memory");
}
__device__ __forceinline__ void stg2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void stg2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
__device__ __forceinline__ void ldg2(const float* p, float& a, float& b) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p));
  a = t.x; b = t.y;
}
__device__ __forceinline__ void ldg2(const double* p, double& a, double& b) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  a = t.x; b = t.y;
}

// Fast arithmetic (reciprocal multiplies + FMA); see the header comment for exactness.
// Every rounding is pinned (explicit fma / non-contractible multiplies): were the scaling by 1/diag an ordinary
// product, nvcc could fuse it into the add of a LATER stage that consumes the value -- in the unmasked fast path
// but not behind the selects of the masked path -- and the last bit of a point would depend on which tile / slab
// it falls in whenever 1/diag is not a power of two (Helmholtz shift, non-dyadic spacings).  Seen on 2 GPUs.
template <typename T>
__device__ __forceinline__ T relax_fast(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T rhs) {
  const T nb = fma(rt + lf, s.ihy2, Strict<T>::mul(up + dn, s.ihx2));
  const T unew = Strict<T>::mul(rhs + nb, s.inv_neg_diag);
  return fma(s.one_minus_omega, uc, Strict<T>::mul(s.omega, unew));
}
template <typename T>
__device__ __forceinline__ T residual_fast(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T f) {
  T t = fma(rt + lf, s.ihy2, Strict<T>::mul(up + dn, s.ihx2));
  t = fma(-uc, s.cc, t);
  return fma(-s.shift, uc, fma(-s.coeff, t, f));  // exact no-op for shift = 0
}

// Isotropic (hx == hy), unrelaxed (omega == 1) specialisation: 5 instead of 8 instructions per point.
// ((a+b)+(c+d))*ih2 equals (a+b)*ih2 + (c+d)*ih2 bit for bit when ih2 is a power of two (scaling by a power
// of two commutes with rounding), and omega = 1 makes the relaxation blend the identity: on dyadic grids without a
// shift this is the reference's value.  Otherwise (Helmholtz shift, non-dyadic h) it is a slightly different --
// equally accurate -- rounding of the same update, and like relax_fast it pins every rounding.
template <typename T>
__device__ __forceinline__ T relax_iso1(const StencilScalars<T>& s, T up, T dn, T rt, T lf, T rhs) {
  const T sum = (up + dn) + (rt + lf);
  return Strict<T>::mul(fma(sum, s.ihx2, rhs), s.inv_neg_diag);
}
template <typename T>
__device__ __forceinline__ T residual_iso(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T f) {
  const T sum = (up + dn) + (rt + lf);
  const T t = fma((T)-4, uc, sum);  // h^2 * lap_h u
  return fma(-s.shift, uc, fma(-s.coeff * s.ihx2, t, f));
}
template <bool SIMPLE, typename T>
__device__ __forceinline__ T relax_sel(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T rhs) {
  if (SIMPLE) return relax_iso1<T>(s, up, dn, rt, lf, rhs);
  return relax_fast<T>(s, uc, up, dn, rt, lf, rhs);
}
// Damped Jacobi point update: the relaxation blend (1-omega)*u_old + omega*u_new keeps the reference's two rounded
// products and rounded sum (smoothers.py:82; no FMA contraction), so with power-of-two spacings the result is
// bit-identical to the strict kernel for any omega.  omega == 1 (SIMPLE) makes the blend the identity.
template <bool SIMPLE, typename T>
__device__ __forceinline__ T relax_jacobi(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T rhs) {
  if (SIMPLE) return relax_iso1<T>(s, up, dn, rt, lf, rhs);
  const T nb = fma(rt + lf, s.ihy2, Strict<T>::mul(up + dn, s.ihx2));
  const T unew = Strict<T>::mul(rhs + nb, s.inv_neg_diag);
  return Strict<T>::add(Strict<T>::mul(s.one_minus_omega, uc), Strict<T>::mul(s.omega, unew));
}
template <bool SIMPLE, typename T>
__device__ __forceinline__ T residual_sel(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T f) {
  if (SIMPLE) return residual_iso<T>(s, uc, up, dn, rt, lf, f);
  return residual_fast<T>(s, uc, up, dn, rt, lf, f);
}

// Variable-coefficient operator  A u = -div(a grad u) + shift*u, arithmetic-mean face coefficients of the nodal field a
// (mg_varcoef.cu states the discretisation; no reference operator exists).  The strict kernel's operation
// order is kept -- two rounded products and their rounded sum per direction, a TRUE division by the per-point diagonal --
// with two exact rewrites: the halving of the face means is folded into the scaling (s = a_nb + a_c is twice the face
// coefficient; multiplying by 0.5 commutes with every rounding), and the divisions by hx^2, hy^2 become multiplications
// by hcx = 0.5/hx^2, hcy = 0.5/hy^2, exact when the spacings are powers of two.  On such grids the fused passes are
// therefore bit-identical to mg_varcoef_smooth_rbgs / mg_varcoef_residual; otherwise they differ by a few ulp.
template <bool NOBLEND, typename T>
__device__ __forceinline__ T relax_var(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T rhs, T ac, T a_up, T a_dn,
                                       T a_rt, T a_lf) {
  using A = Strict<T>;
  const T se = A::add(a_up, ac), sw = A::add(a_dn, ac), sn = A::add(a_rt, ac), ss = A::add(a_lf, ac);
  const T x = A::add(A::mul(se, up), A::mul(sw, dn));
  const T y = A::add(A::mul(sn, rt), A::mul(ss, lf));
  const T nb = A::add(A::mul(x, s.ihx2), A::mul(y, s.ihy2));  // VARCOEF scalars: ihx2 = 0.5/hx^2, ihy2 = 0.5/hy^2
  const T diag = A::add(A::add(A::mul(A::add(se, sw), s.ihx2), A::mul(A::add(sn, ss), s.ihy2)), s.shift);
  const T unew = A::div(A::add(rhs, nb), diag);
  if (NOBLEND) return unew;  // omega == 1
  return A::add(A::mul(s.one_minus_omega, uc), A::mul(s.omega, unew));
}
template <typename T>
__device__ __forceinline__ T residual_var(const StencilScalars<T>& s, T uc, T up, T dn, T rt, T lf, T f, T ac, T a_up, T a_dn,
                                          T a_rt, T a_lf) {
  using A = Strict<T>;
  const T se = A::add(a_up, ac), sw = A::add(a_dn, ac), sn = A::add(a_rt, ac), ss = A::add(a_lf, ac);
  const T x = A::mul(A::add(A::mul(se, A::sub(uc, up)), A::mul(sw, A::sub(uc, dn))), s.ihx2);
  const T y = A::mul(A::add(A::mul(sn, A::sub(uc, rt)), A::mul(ss, A::sub(uc, lf))), s.ihy2);
  return A::sub(f, A::add(A::add(x, y), A::mul(s.shift, uc)));
}

// ---------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int NU, int FRONT, int BACK, int LOADER, bool SIMPLE, int SMOOTH, int WARPS, int NSTAGE, int RB,
          bool VARCOEF = false>
__global__ void __launch_bounds__(WARPS * 32)
    rbgs_stream_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_f,
                       const __grid_constant__ CUtensorMap map_e, const __grid_constant__ CUtensorMap map_a,
                       const PassParams p, const StencilScalars<T> sc) {
  constexpr int NS = num_stages(SMOOTH, NU);  // pipeline stages; the row loaded NS steps ago is final
  static_assert(!VARCOEF || (SMOOTH == SMOOTH_RBGS && LOADER == LOADER_TMA),
                "variable coefficients: red-black GS, TMA-staged only");
  using G = Geometry<NS, BACK>;
  static_assert(RB % 2 == 0, "row parity must be static inside a box");
  static_assert(FRONT != FRONT_ADDFINE || sizeof(T) == 8, "ADDFINE adds an fp32 correction to an fp64 iterate");
  static_assert(BACK != BACK_RESID || sizeof(T) == 8, "RESID rounds an fp64 residual to fp32");
  constexpr int WR = G::WR, FR = G::FR;
  constexpr bool HAS_BACK = BACK != BACK_NONE;
  constexpr bool HAS_NORM = BACK == BACK_NORM || BACK == BACK_RESID;
  constexpr uint32_t ROW_BYTES = STRIP * sizeof(T);
  constexpr uint32_t BOX_BYTES = RB * ROW_BYTES;                            // one T array, one stage
  constexpr uint32_t EBOX_BYTES = FRONT == FRONT_ADDFINE ? RB * STRIP * 4 : 0;  // fp32 correction box
  // coarse rows of a prolongation pass: staged by TMA for fp32 (the __ldg path exposes ~6 long-scoreboard stalls per
  // issue slot, ncu r01), read directly for fp64 / cp.async
  constexpr bool STAGE_COARSE = StageCoarse<T, FRONT, BACK, LOADER>::value;
  constexpr int CALIGN = 16 / (int)sizeof(T) - 1;  // box start rounded down to 16 bytes
  constexpr uint32_t CBOX_TX = STAGE_COARSE ? (RB / 2 + 1) * COARSE_BOX_W * sizeof(T) : 0;  // bytes TMA delivers
  constexpr uint32_t CBOX_BYTES = (CBOX_TX + 127u) & ~127u;                                     // ring slot (128-B aligned)
  constexpr uint32_t ABOX_BYTES = VARCOEF ? BOX_BYTES : 0;  // nodal coefficient rows ride the same ring (+1 word / point)
  constexpr uint32_t ABOX_OFF = 2 * BOX_BYTES + EBOX_BYTES + CBOX_BYTES;
  constexpr uint32_t STAGE_BYTES = ABOX_OFF + ABOX_BYTES;

  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[WARPS][NSTAGE];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x * WARPS + warp;
  unsigned char* ring = smem + (size_t)warp * NSTAGE * STAGE_BYTES;  // [stage][u box | f box | e box]
  const uint32_t ring_a = smem_u32(ring);                            // the same, as a shared-window address

  if (LOADER == LOADER_TMA) {
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[warp][s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
  if (strip >= p.nstrips) {  // warp-uniform; warps never synchronise with each other
    if (HAS_NORM && lane == 0) p.partials[(size_t)blockIdx.y * (gridDim.x * WARPS) + strip] = 0.0;
    return;
  }

  const int nx = p.nx, ny = p.ny;
  const int g0 = strip * G::STRIDE - 4;  // global column of local column 0 (multiple of 4)
  const int jbase = g0 + lane * LANE_V;  // global column of this lane's element 0
  const int I0 = blockIdx.y * p.rows_per_tile;
  const int I1 = min(I0 + p.rows_per_tile, nx);  // owned rows [I0, I1)
  const int i_begin = I0 - G::ROW_LEAD;          // even
  const int i_last = I1 - 1 + G::ROW_TAIL;
  const int nbox = (i_last - i_begin + 1 + RB - 1) / RB;
  const bool u_zero = p.u_zero != 0;
  // a strip whose 128 columns are all interior needs no column masks
  const bool strip_interior = (g0 >= 1) && (g0 + STRIP - 1 <= ny - 2);

  // per-element masks
  uint32_t upd = 0, dom = 0, own = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = jbase + e, x = lane * LANE_V + e;
    const bool d = (j >= 0 && j < ny);
    const bool o = d && x <= G::OWN_HI && (x >= G::OWN_LO || (strip == 0 && x >= 4));
    upd |= (j >= 1 && j <= ny - 2) ? (1u << e) : 0u;
    dom |= d ? (1u << e) : 0u;
    own |= o ? (1u << e) : 0u;
  }

  const bool own_all = own == 0xFu, own_lo = own == 0x3u, own_hi = own == 0xCu;  // per-lane constants

  // ---- loader ----------------------------------------------------------------------------------
  auto issue_box = [&](int box) {
    const int stage = box % NSTAGE;
    unsigned char* dst_u = ring + (size_t)stage * STAGE_BYTES;
    unsigned char* dst_f = dst_u + BOX_BYTES;
    unsigned char* dst_e = dst_f + BOX_BYTES;
    const int row0 = i_begin + box * RB;
    if (LOADER == LOADER_TMA) {
      if (lane == 0) {
        mbar_expect_tx(&full_bar[warp][stage], (u_zero ? 0u : BOX_BYTES) + BOX_BYTES + EBOX_BYTES + CBOX_TX + ABOX_BYTES);
        if (!u_zero) tma_load_2d(dst_u, &map_u, g0, row0, &full_bar[warp][stage]);
        if (VARCOEF) tma_load_2d(dst_u + ABOX_OFF, &map_a, g0, row0, &full_bar[warp][stage]);
        tma_load_2d(dst_f, &map_f, g0, row0, &full_bar[warp][stage]);
        if (FRONT == FRONT_ADDFINE) tma_load_2d(dst_e, &map_e, g0, row0, &full_bar[warp][stage]);
        // coarse rows row0/2 .. row0/2 + RB/2, coarse columns g0/2 .. (out-of-range parts arrive as zeros)
        // (the box starts at a multiple of 4 columns = 16 bytes: sm_100a faulted on 8-byte aligned box starts)
        if (STAGE_COARSE) tma_load_2d(dst_e, &map_e, (g0 >> 1) & ~CALIGN, row0 >> 1, &full_bar[warp][stage]);
      }
    } else {
      // per lane: its own 4 elements of every row of the box; zero-fill outside the domain
      int nvalid = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) nvalid += (jbase + e >= 0 && jbase + e < ny) ? 1 : 0;  // a prefix, or empty
      const bool colok = (jbase >= 0) && nvalid > 0;
      const int jc = colok ? jbase : 0;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int row = row0 + r;
        const bool ok = colok && row >= 0 && row < nx;
        const int rc = ok ? row : 0;
        const int bytes = ok ? nvalid * (int)sizeof(T) : 0;
        const T* su = reinterpret_cast<const T*>(p.u_in) + (int64_t)rc * p.ld_in + jc;
        const T* sf = reinterpret_cast<const T*>(p.f) + (int64_t)rc * p.ld_f + jc;
        unsigned char* du = dst_u + r * ROW_BYTES + lane * LANE_V * sizeof(T);
        unsigned char* df = dst_f + r * ROW_BYTES + lane * LANE_V * sizeof(T);
        if (sizeof(T) == 4) {
          if (!u_zero) cp_async16(du, su, bytes);
          cp_async16(df, sf, bytes);
        } else {
          if (!u_zero) {
            cp_async16(du, su, min(bytes, 16));
            cp_async16(du + 16, su + 2, max(bytes - 16, 0));
          }
          cp_async16(df, sf, min(bytes, 16));
          cp_async16(df + 16, sf + 2, max(bytes - 16, 0));
        }
        if (FRONT == FRONT_ADDFINE)
          cp_async16(dst_e + r * STRIP * 4 + lane * LANE_V * 4, p.fine_in + (int64_t)rc * p.ld_fi + jc,
                     ok ? nvalid * 4 : 0);
      }
    }
  };

  // prologue: fill the ring
#pragma unroll
  for (int b = 0; b < NSTAGE; ++b) {
    if (b < nbox) issue_box(b);
    if (LOADER == LOADER_CPASYNC) cp_async_commit();  // one group per slot, even when empty
  }

  // ---- state -----------------------------------------------------------------------------------
  T w[WR][4];   // w[a] = row loaded a steps ago (age a)
  T fr[FR][4];  // matching right-hand sides
#pragma unroll
  for (int a = 0; a < WR; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) w[a][e] = (T)0;
#pragma unroll
  for (int a = 0; a < FR; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) fr[a][e] = (T)0;
  // Jacobi: stage s also needs the PREVIOUS iterate of the older neighbour row, which stage s itself overwrote one
  // step earlier; pj[s] keeps that row (version s-1 of the row of age s+1)
  T pj[NS + 1][4];
#pragma unroll
  for (int a = 0; a <= NS; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) pj[a][e] = (T)0;
  // variable coefficients: the same row window for the nodal field a; [4] / [5] = the left / right neighbour column of
  // the lane's elements 0 / 3 (one shuffle each when the row arrives: a does not change between the stages)
  T ac[VARCOEF ? WR : 1][6];
#pragma unroll
  for (int a = 0; a < (VARCOEF ? WR : 1); ++a)
#pragma unroll
    for (int e = 0; e < 6; ++e) ac[a][e] = (T)0;
  T rr[3][4];  // residual rows (BACK_RESTRICT): rr[0] newest
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) rr[a][e] = (T)0;
  double acc = 0.0;  // sum of squared residuals

  // coarse correction rows (FRONT_PROLONG): c0 = coarse row floor(i/2), c1 = the next one; 3 columns each
  T c0[3] = {(T)0, (T)0, (T)0}, c1[3] = {(T)0, (T)0, (T)0};
  const int jc0 = (g0 >> 1) + 2 * lane;  // coarse column of element 0 (g0 is a multiple of 4, may be -4)
  auto load_coarse_row = [&](int ic, T(&c)[3], bool checked) {
    T a = (T)0, b = (T)0, d = (T)0;
    const T* row = reinterpret_cast<const T*>(p.coarse_in) + (int64_t)ic * p.ld_ci;
    if (!checked) {
      ldg2(row + jc0, a, b);
      if (lane == 31) d = __ldg(row + jc0 + 2);
    } else if (ic >= 0 && ic < p.nxc) {
      if (jc0 >= 0 && jc0 + 1 < p.nyc) {
        ldg2(row + jc0, a, b);
      } else {
        if (jc0 >= 0 && jc0 < p.nyc) a = __ldg(row + jc0);
        if (jc0 + 1 >= 0 && jc0 + 1 < p.nyc) b = __ldg(row + jc0 + 1);
      }
      if (lane == 31 && jc0 + 2 >= 0 && jc0 + 2 < p.nyc) d = __ldg(row + jc0 + 2);
    }
    const T nxt = shfl_dn1(a);
    c[0] = a; c[1] = b; c[2] = (lane == 31) ? d : nxt;
  };
  if (FRONT == FRONT_PROLONG && !STAGE_COARSE) {
    load_coarse_row(i_begin >> 1, c0, true);  // i_begin is even; >> floors for negatives
    load_coarse_row((i_begin >> 1) + 1, c1, true);
  }
  // staged variant: row `rel` (0 .. RB/2) of the current box's coarse slab, columns 2*lane .. 2*lane + 2
  const uint32_t coarse_lane_off = (uint32_t)((((g0 >> 1) & CALIGN) + 2 * lane) * (int)sizeof(T));
  auto read_coarse_row = [&](uint32_t sc_box, int rel, T(&c)[3]) {
    const uint32_t row = sc_box + (uint32_t)(rel * COARSE_BOX_W * (int)sizeof(T)) + coarse_lane_off;
    lds2(row, c[0], c[1]);
    lds1(row + 2u * (uint32_t)sizeof(T), c[2]);
  };

  T* const uout = reinterpret_cast<T*>(p.u_out);
  T* const cout = reinterpret_cast<T*>(p.coarse_out);
  const bool store_u = p.store_u != 0;

  // ---- one box of RB rows.  MASKED = false is the interior fast path: every row and column the box
  //      touches is an interior point, so there are no boundary tests, selects or divergent branches. ----
  // su / sf / se: shared-window addresses of this lane's 4 elements of the box's first u / f / fp32-correction row;
  // sc_box: of the box's coarse slab (STAGE_COARSE)
  auto process_box = [&](auto masked_tag, const int ib, const uint32_t su, const uint32_t sf, const uint32_t se,
                         const uint32_t sc_box, const uint32_t sa) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    if (STAGE_COARSE) {
      read_coarse_row(sc_box, 0, c0);
      read_coarse_row(sc_box, 1, c1);
    }
    // rows stored by this box: ib - NS ... ib + RB - 1 - NS; all owned by the tile?  (uniform)
    const bool rows_owned = (ib - NS >= I0) && (ib + RB - 1 - NS < I1);
    T* orow = uout + (int64_t)(ib - NS) * p.ld_out + jbase;  // running output row pointer
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int i = ib + k;  // newest row; parity of i == parity of k
      const int kpar = k & 1;

      // (0) shift the windows by one row (register renaming after unrolling)
#pragma unroll
      for (int a = WR - 1; a > 0; --a)
#pragma unroll
        for (int e = 0; e < 4; ++e) w[a][e] = w[a - 1][e];
#pragma unroll
      for (int a = FR - 1; a > 0; --a)
#pragma unroll
        for (int e = 0; e < 4; ++e) fr[a][e] = fr[a - 1][e];
      if (VARCOEF) {
#pragma unroll
        for (int a = WR - 1; a > 0; --a)
#pragma unroll
          for (int e = 0; e < 6; ++e) ac[a][e] = ac[a - 1][e];
      }

      // (1) newest row from the ring
      if (u_zero) {
#pragma unroll
        for (int e = 0; e < 4; ++e) w[0][e] = (T)0;
      } else {
        lds4(su + (uint32_t)(k * STRIP * (int)sizeof(T)), w[0]);
      }
      lds4(sf + (uint32_t)(k * STRIP * (int)sizeof(T)), fr[0]);
      if (VARCOEF) {
        T a4[4];
        lds4(sa + (uint32_t)(k * STRIP * (int)sizeof(T)), a4);
        ac[0][0] = a4[0]; ac[0][1] = a4[1]; ac[0][2] = a4[2]; ac[0][3] = a4[3];
        ac[0][4] = shfl_up1(a4[3]);
        ac[0][5] = shfl_dn1(a4[0]);
      }

      // (1a) FRONT_ADDFINE: u += (T)e  (rows/columns outside the domain hold zeros in both arrays)
      if (FRONT == FRONT_ADDFINE) {
        float ev[4];
        lds4(se + (uint32_t)(k * STRIP * 4), ev);
        w[0][0] += (T)ev[0]; w[0][1] += (T)ev[1]; w[0][2] += (T)ev[2]; w[0][3] += (T)ev[3];
      }

      // (1b) FRONT_PROLONG: u += bilinear prolongation of the coarse correction (transfer.py:234-267 semantics)
      if (FRONT == FRONT_PROLONG) {
        if (!MASKED || (i >= 0 && i < nx)) {
          T add[4];
          if (kpar == 0) {  // even fine row: coarse row i/2
            const bool last = MASKED && (i == nx - 1);
            add[0] = c0[0];
            add[2] = c0[1];
            add[1] = last ? (T)0 : (T)0.5 * (c0[0] + c0[1]);
            add[3] = last ? (T)0 : (T)0.5 * (c0[1] + c0[2]);
          } else {  // odd fine row: coarse rows (i-1)/2 and (i+1)/2
            add[0] = (MASKED && jbase == ny - 1) ? (T)0 : (T)0.5 * (c0[0] + c1[0]);
            add[2] = (MASKED && jbase + 2 == ny - 1) ? (T)0 : (T)0.5 * (c0[1] + c1[1]);
            add[1] = (T)0.25 * (((c0[0] + c0[1]) + c1[0]) + c1[1]);
            add[3] = (T)0.25 * (((c0[1] + c0[2]) + c1[1]) + c1[2]);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) w[0][e] = (!MASKED || ((dom >> e) & 1u)) ? w[0][e] + add[e] : w[0][e];
        }
        if (kpar == 1) {  // moving to the next coarse row pair
#pragma unroll
          for (int e = 0; e < 3; ++e) c0[e] = c1[e];
          if (STAGE_COARSE) {
            if (k + 1 < RB) read_coarse_row(sc_box, (k + 1) / 2 + 1, c1);  // the next box reloads both rows
          } else {
            load_coarse_row(((i + 1) >> 1) + 1, c1, MASKED);
          }
        }
      }

      // (2) the NS stages, stage s on the row of age s
#pragma unroll
      for (int s = 1; s <= NS; ++s) {
        const int q = i - s;
        if (SMOOTH == SMOOTH_JACOBI) {
          // damped Jacobi (smoothers.py:41-86): every point from the previous iterate; the newer neighbour row (age
          // s-1) was brought to iterate s-1 by stage s-1 in this step, the older one is kept in pj[s]
          T old[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) old[e] = w[s][e];
          if (!MASKED || (q >= 1 && q <= nx - 2)) {  // warp-uniform
            const T lfx = shfl_up1(old[3]);
            const T rtx = shfl_dn1(old[0]);
            const T n0 = relax_jacobi<SIMPLE, T>(sc, old[0], w[s - 1][0], pj[s][0], old[1], lfx, fr[s][0]);
            const T n1 = relax_jacobi<SIMPLE, T>(sc, old[1], w[s - 1][1], pj[s][1], old[2], old[0], fr[s][1]);
            const T n2 = relax_jacobi<SIMPLE, T>(sc, old[2], w[s - 1][2], pj[s][2], old[3], old[1], fr[s][2]);
            const T n3 = relax_jacobi<SIMPLE, T>(sc, old[3], w[s - 1][3], pj[s][3], rtx, old[2], fr[s][3]);
            w[s][0] = (!MASKED || (upd & 1u)) ? n0 : old[0];
            w[s][1] = (!MASKED || (upd & 2u)) ? n1 : old[1];
            w[s][2] = (!MASKED || (upd & 4u)) ? n2 : old[2];
            w[s][3] = (!MASKED || (upd & 8u)) ? n3 : old[3];
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) pj[s][e] = old[e];
        } else if (!MASKED || (q >= 1 && q <= nx - 2)) {  // warp-uniform
          // colour of stage s is (s-1)&1 (red = (row+col) even first); col parity == element parity
          const int e0 = (kpar + s + ((s - 1) & 1)) & 1;  // first updated element, compile-time after unrolling
          if (e0 == 0) {
            const T lfx = shfl_up1(w[s][3]);
            T n0, n2;
            if (VARCOEF) {
              constexpr int sa_ = VARCOEF ? 1 : 0;  // keeps the indices inside ac[1][6] in the constant-coefficient kernels
              n0 = relax_var<SIMPLE, T>(sc, w[s][0], w[s - 1][0], w[s + 1][0], w[s][1], lfx, fr[s][0], ac[s * sa_][0],
                                        ac[(s - 1) * sa_][0], ac[(s + 1) * sa_][0], ac[s * sa_][1], ac[s * sa_][4]);
              n2 = relax_var<SIMPLE, T>(sc, w[s][2], w[s - 1][2], w[s + 1][2], w[s][3], w[s][1], fr[s][2], ac[s * sa_][2],
                                        ac[(s - 1) * sa_][2], ac[(s + 1) * sa_][2], ac[s * sa_][3], ac[s * sa_][1]);
            } else {
              n0 = relax_sel<SIMPLE, T>(sc, w[s][0], w[s - 1][0], w[s + 1][0], w[s][1], lfx, fr[s][0]);
              n2 = relax_sel<SIMPLE, T>(sc, w[s][2], w[s - 1][2], w[s + 1][2], w[s][3], w[s][1], fr[s][2]);
            }
            w[s][0] = (!MASKED || (upd & 1u)) ? n0 : w[s][0];
            w[s][2] = (!MASKED || (upd & 4u)) ? n2 : w[s][2];
          } else {
            const T rtx = shfl_dn1(w[s][0]);
            T n1, n3;
            if (VARCOEF) {
              constexpr int sa_ = VARCOEF ? 1 : 0;
              n1 = relax_var<SIMPLE, T>(sc, w[s][1], w[s - 1][1], w[s + 1][1], w[s][2], w[s][0], fr[s][1], ac[s * sa_][1],
                                        ac[(s - 1) * sa_][1], ac[(s + 1) * sa_][1], ac[s * sa_][2], ac[s * sa_][0]);
              n3 = relax_var<SIMPLE, T>(sc, w[s][3], w[s - 1][3], w[s + 1][3], rtx, w[s][2], fr[s][3], ac[s * sa_][3],
                                        ac[(s - 1) * sa_][3], ac[(s + 1) * sa_][3], ac[s * sa_][5], ac[s * sa_][2]);
            } else {
              n1 = relax_sel<SIMPLE, T>(sc, w[s][1], w[s - 1][1], w[s + 1][1], w[s][2], w[s][0], fr[s][1]);
              n3 = relax_sel<SIMPLE, T>(sc, w[s][3], w[s - 1][3], w[s + 1][3], rtx, w[s][2], fr[s][3]);
            }
            w[s][1] = (!MASKED || (upd & 2u)) ? n1 : w[s][1];
            w[s][3] = (!MASKED || (upd & 8u)) ? n3 : w[s][3];
          }
        }
      }

      // (3) the row of age NS is final: store the owned part (predicated stores, no divergent branch)
      {
        const int qf = i - NS;
        const bool row_ok = store_u && (rows_owned || (qf >= I0 && qf < I1));
        T* dst = orow;
        orow += p.ld_out;
        if (!MASKED) {  // interior strip: ownership changes only at even columns; three predicated stores, no branch
          stg4_if(row_ok && own_all, dst, w[NS][0], w[NS][1], w[NS][2], w[NS][3]);
          stg2_if(row_ok && own_lo, dst, w[NS][0], w[NS][1]);
          stg2_if(row_ok && own_hi, dst + 2, w[NS][2], w[NS][3]);
        } else if (row_ok && own != 0u) {
          if (own == 0xFu) {
            stg4(dst, w[NS]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if ((own >> e) & 1u) dst[e] = w[NS][e];
          }
        }
      }

      // (4) BACK: residual of the row of age NS+1 (its neighbours are final)
      if (HAS_BACK) {
        constexpr int A = NS + 1;
        const int q2 = i - A;
        T r[4] = {(T)0, (T)0, (T)0, (T)0};
        if (!MASKED || (q2 >= 0 && q2 < nx)) {
          if (!MASKED || (q2 >= 1 && q2 <= nx - 2)) {
            const T lfx = shfl_up1(w[A][3]);
            const T rtx = shfl_dn1(w[A][0]);
            T r0, r1, r2, r3;
            if (VARCOEF) {
              constexpr int B0 = VARCOEF ? A : 0, Bm = VARCOEF ? A - 1 : 0, Bp = VARCOEF ? A + 1 : 0;
              r0 = residual_var<T>(sc, w[A][0], w[A - 1][0], w[A + 1][0], w[A][1], lfx, fr[A][0], ac[B0][0], ac[Bm][0], ac[Bp][0],
                                   ac[B0][1], ac[B0][4]);
              r1 = residual_var<T>(sc, w[A][1], w[A - 1][1], w[A + 1][1], w[A][2], w[A][0], fr[A][1], ac[B0][1], ac[Bm][1],
                                   ac[Bp][1], ac[B0][2], ac[B0][0]);
              r2 = residual_var<T>(sc, w[A][2], w[A - 1][2], w[A + 1][2], w[A][3], w[A][1], fr[A][2], ac[B0][2], ac[Bm][2],
                                   ac[Bp][2], ac[B0][3], ac[B0][1]);
              r3 = residual_var<T>(sc, w[A][3], w[A - 1][3], w[A + 1][3], rtx, w[A][2], fr[A][3], ac[B0][3], ac[Bm][3], ac[Bp][3],
                                   ac[B0][5], ac[B0][2]);
            } else {
              r0 = residual_sel<SIMPLE, T>(sc, w[A][0], w[A - 1][0], w[A + 1][0], w[A][1], lfx, fr[A][0]);
              r1 = residual_sel<SIMPLE, T>(sc, w[A][1], w[A - 1][1], w[A + 1][1], w[A][2], w[A][0], fr[A][1]);
              r2 = residual_sel<SIMPLE, T>(sc, w[A][2], w[A - 1][2], w[A + 1][2], w[A][3], w[A][1], fr[A][2]);
              r3 = residual_sel<SIMPLE, T>(sc, w[A][3], w[A - 1][3], w[A + 1][3], rtx, w[A][2], fr[A][3]);
            }
            r[0] = (!MASKED || (upd & 1u)) ? r0 : fr[A][0];
            r[1] = (!MASKED || (upd & 2u)) ? r1 : fr[A][1];
            r[2] = (!MASKED || (upd & 4u)) ? r2 : fr[A][2];
            r[3] = (!MASKED || (upd & 8u)) ? r3 : fr[A][3];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = fr[A][e];  // boundary rows: r = f (laplacian.py:64,117)
          }
        }
        if (HAS_NORM) {
          if (q2 >= I0 && q2 < I1) {
            if (q2 >= p.norm_row_lo && q2 < p.norm_row_hi) {
              // squares formed in T (like NumPy's field**2), the row's four summed in T, one fp64 add per row
              T rowsum;
              if (!MASKED && own == 0xFu) {
                rowsum = (r[0] * r[0] + r[1] * r[1]) + (r[2] * r[2] + r[3] * r[3]);
              } else {
                const T m0 = (own & 1u) ? r[0] : (T)0, m1 = (own & 2u) ? r[1] : (T)0;
                const T m2 = (own & 4u) ? r[2] : (T)0, m3 = (own & 8u) ? r[3] : (T)0;
                rowsum = (m0 * m0 + m1 * m1) + (m2 * m2 + m3 * m3);
              }
              acc += (double)rowsum;
            }
            if (BACK == BACK_RESID && own != 0u) {
              float* dst = p.resid_out + (int64_t)q2 * p.ld_ro + jbase;
              if (own == 0xFu) {
                *reinterpret_cast<float4*>(dst) = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if ((own >> e) & 1u) dst[e] = (float)r[e];
              }
            }
          }
        } else {  // BACK_RESTRICT
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            rr[2][e] = rr[1][e];
            rr[1][e] = rr[0][e];
            rr[0][e] = r[e];
          }
          if (kpar == (NS & 1)) {  // q2 = i - NS - 1 is odd: rows q2-2, q2-1 (centre, even), q2 are complete
            const int fi = q2 - 1;  // fine centre row
            const int ic = fi >> 1;
            if (fi >= I0 && fi < I1) {  // owned coarse row (warp-uniform); I0 >= 0, I1 <= nx
              const T l2 = shfl_up1(rr[2][3]), l1 = shfl_up1(rr[1][3]), l0 = shfl_up1(rr[0][3]);
              const bool brow = MASKED && (ic == 0 || ic == p.nxc - 1);
              // element 0 -> coarse column jc0, element 2 -> coarse column jc0 + 1
              T v0, v1;
              {
                const T corners = ((l2 + rr[2][1]) + l0) + rr[0][1];
                const T edges = ((rr[2][0] + rr[0][0]) + l1) + rr[1][1];
                v0 = ((T)0.0625 * corners + (T)0.125 * edges) + (T)0.25 * rr[1][0];
                const bool b = MASKED && (brow || jc0 == 0 || jc0 == p.nyc - 1);
                v0 = b ? rr[1][0] : v0;
              }
              {
                const T corners = ((rr[2][1] + rr[2][3]) + rr[0][1]) + rr[0][3];
                const T edges = ((rr[2][2] + rr[0][2]) + rr[1][1]) + rr[1][3];
                v1 = ((T)0.0625 * corners + (T)0.125 * edges) + (T)0.25 * rr[1][2];
                const bool b = MASKED && (brow || jc0 + 1 == 0 || jc0 + 1 == p.nyc - 1);
                v1 = b ? rr[1][2] : v1;
              }
              const bool o0 = (own & 1u) != 0u, o1 = (own & 4u) != 0u;
              T* dst = cout + (int64_t)ic * p.ld_co + jc0;
              stg2_if(o0 && o1, dst, v0, v1);
              stg1_if(o0 && !o1, dst, v0);
              stg1_if(o1 && !o0, dst + 1, v1);
            }
          }
        }
      }
    }  // rows of the box
  };

  // ---- main loop over boxes ----------------------------------------------------------------------
  for (int box = 0; box < nbox; ++box) {
    const int stage = box % NSTAGE;
    if (LOADER == LOADER_TMA) {
      mbar_wait(&full_bar[warp][stage], (uint32_t)((box / NSTAGE) & 1));
    } else {
      cp_async_wait<NSTAGE - 1>();
      __syncwarp();
    }
    const uint32_t sbox = ring_a + (uint32_t)stage * STAGE_BYTES;
    const uint32_t su = sbox + (uint32_t)(lane * LANE_V * (int)sizeof(T));
    const uint32_t sf = su + BOX_BYTES;
    const uint32_t sc_box = sbox + 2 * BOX_BYTES;              // coarse slab (STAGE_COARSE) ...
    const uint32_t se = sc_box + (uint32_t)(lane * LANE_V * 4);  // ... or fp32 correction rows (FRONT_ADDFINE)
    const uint32_t sa = su + ABOX_OFF;                           // nodal coefficient rows (VARCOEF)
    const int ib = i_begin + box * RB;
    // interior fast path: every row evaluated in this box (oldest: ib - NS - 1 with a BACK stage) and the
    // newest row ib + RB - 1 are interior rows, and the strip has no boundary column
    // (with restriction also the centre row q2 - 1 of the oldest coarse row, hence 3 instead of 1)
    constexpr int OLDEST = NS + (BACK == BACK_RESTRICT ? 3 : (HAS_BACK ? 1 : 0));
    const bool fast = strip_interior && (ib - OLDEST >= 1) && (ib + RB - 1 <= nx - 2);
    if (fast) process_box(FalseTag{}, ib, su, sf, se, sc_box, sa);
    else process_box(TrueTag{}, ib, su, sf, se, sc_box, sa);

    // refill this stage with box + NSTAGE
    __syncwarp();
    if (box + NSTAGE < nbox) issue_box(box + NSTAGE);
    if (LOADER == LOADER_CPASYNC) cp_async_commit();
  }

  if (HAS_NORM) {
    acc = warp_sum(acc);
    if (lane == 0) p.partials[(size_t)blockIdx.y * (gridDim.x * WARPS) + strip] = acc;
  }
}

}  // namespace stream
}  // namespace mg
