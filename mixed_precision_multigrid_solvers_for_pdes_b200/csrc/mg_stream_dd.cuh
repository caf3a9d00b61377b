// libmgb200: the fused DEFECT + DOWN pass of the mixed-precision refinement cycle (sm_100a).
//
// One refinement cycle on level 0 used to be three passes over HBM:
//   defect pass   u64 += e32 ; r32 = fp32(f64 - A u64) ; sum r64^2                     32 B / point
//   down pass     e = 0 ; 2 RB-GS sweeps on A e = r32 ; f_c = R(r32 - A e)             9 B / point  (reads r32 again)
//   up pass       e += P e_c ; 2 RB-GS sweeps                                          13 B / point
// This kernel does the first two in ONE pass: the fp32 residual row leaves for HBM (the up pass needs it as its
// right-hand side) but is consumed by the smoothing pipeline straight from registers: 37 B / point instead of 41, one
// launch less, and the issue-bound smoothing instructions (75 % issue utilisation on their own) run in the shadow of the
// HBM-bound defect traffic (measured at 16385^2: 1.70-1.73 ms against 1.37 + 0.49, DRAM traffic at 0.98 of the copy peak;
// profiles/r02_ncu_dd_16385.md).  Same decomposition as rbgs_stream_kernel (mg_stream.cuh): every warp owns a strip of 128
// columns and streams down a tile of rows, TMA boxes of RB rows through a warp-private ring, register windows, recomputed
// halos: 8 columns per side (1 for the fp64 residual + 4 half-sweeps + 2 for the restriction, rounded up to the 4-element
// vectors), 6 + 6 rows (the dependence cone exactly, see DDGeometry).
//
// Per arriving row i (per lane 4 contiguous elements):
//   u_new(i)  = u(i) + (double) e(i)                               -> stored (fp64)
//   r64(i-1)  = f(i-1) - A u_new(i-1)   [needs u_new(i-2 .. i)]    -> sum of squares (fp64), r32(i-1) stored (fp32)
//   fp32 pipeline on i' = i-1: e'(i') = 0, rhs = r32(i');  half-sweep stage s on row i'-s;  row i'-4 final -> stored
//   residual of e' on row i'-5, full weighting centred on row i'-6 when that row is even     -> coarse rhs stored
// Every arithmetic expression is the one of the separate passes (residual_sel / relax_sel, reference summation order in
// the restriction), so u_new, r32, e' and f_c are BIT-IDENTICAL to mg_vc_defect_pass followed by mg_vc_pass
// (tests/test_gpu_fused_parity.py); only the partial sums of the norm are grouped differently (other strip width).
#pragma once
#include "mg_stream.cuh"

namespace mg {
namespace stream {

struct DDGeometry {
  static constexpr int NS = 4;          // two red-black sweeps = four half-sweep stages
  static constexpr int H = 8;           // halo columns per side
  static constexpr int OWN_LO = 8, OWN_HI = STRIP - 1 - OWN_LO, STRIDE = OWN_HI - OWN_LO + 1;  // 112 owned columns
  // Rows a tile streams besides its own [I0, I1).  Dependence cone: the restricted residual centred on row I0 reads
  // residual rows >= I0-1, those read e' rows >= I0-2, four half-sweeps from a ZERO iterate reach three rows further
  // into the right-hand side (r32 rows >= I0-5), and r32(I0-5) reads u row I0-6.  Downwards the loop emits the
  // restriction centred on row i-7 when it loads row i: centres <= I1-2 in tiles that end on an even row (all but
  // the last, which ends at the odd nx and needs one row more; rows >= nx are never loaded anyway).
  static constexpr int ROW_LEAD = 6, ROW_TAIL = 6, ROW_TAIL_LAST = 8;
  static constexpr int WR = NS + 3;     // e' window: ages 0 .. NS+2 (residual stage reads ages NS .. NS+2)
  static constexpr int FR = NS + 2;     // r32 window: ages 0 .. NS+1
};

struct DDParams {
  const double* u_in;    // fp64 iterate (not read when u_zero)
  double* u_out;         // u_in + e_in (written when has_e)
  const double* f;       // fp64 right-hand side
  const float* e_in;     // fp32 correction of the previous cycle (read when has_e)
  float* r_out;          // fp32 residual of the new iterate: right-hand side of the error equation
  float* e_out;          // fp32 error iterate after the pre-smoothing sweeps
  float* coarse_out;     // restricted residual of the error equation
  double* partials;      // one double per warp: sum of the squared fp64 residual
  int nx, ny, nxc, nyc;
  int64_t ld_u, ld_uo, ld_f, ld_e, ld_r, ld_eo, ld_co;
  int u_zero, has_e;
  int norm_row_lo, norm_row_hi;
  int rows_per_tile, nstrips;
};

template <bool SIMPLE, int WARPS, int NSTAGE, int RB, int MINB = 1>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    defect_down_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_f,
                       const __grid_constant__ CUtensorMap map_e, const DDParams p, const StencilScalars<double> sd,
                       const StencilScalars<float> sf) {
  using G = DDGeometry;
  constexpr int NS = G::NS, WR = G::WR, FR = G::FR;
  static_assert(RB % 2 == 0, "row parity must be static inside a box");
  constexpr uint32_t BOX64 = RB * STRIP * 8, BOX32 = RB * STRIP * 4;
  constexpr uint32_t STAGE_BYTES = 2 * BOX64 + BOX32;  // [u64 | f64 | e32]

  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[WARPS][NSTAGE];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x * WARPS + warp;
  unsigned char* ring = smem + (size_t)warp * NSTAGE * STAGE_BYTES;
  const uint32_t ring_a = smem_u32(ring);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (strip >= p.nstrips) {  // warp-uniform; warps never synchronise with each other
    if (lane == 0) p.partials[(size_t)blockIdx.y * (gridDim.x * WARPS) + strip] = 0.0;
    return;
  }

  const int nx = p.nx, ny = p.ny;
  const int g0 = strip * G::STRIDE - 4;  // global column of local column 0 (multiple of 4)
  const int jbase = g0 + lane * LANE_V;
  const int I0 = blockIdx.y * p.rows_per_tile;
  const int I1 = min(I0 + p.rows_per_tile, nx);
  const int i_begin = I0 - G::ROW_LEAD;  // even
  const int i_last = I1 - 1 + ((I1 & 1) ? G::ROW_TAIL_LAST : G::ROW_TAIL);
  const int nbox = (i_last - i_begin + 1 + RB - 1) / RB;
  const bool u_zero = p.u_zero != 0, has_e = p.has_e != 0;
  const bool strip_interior = (g0 >= 1) && (g0 + STRIP - 1 <= ny - 2);

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
  const bool own_all = own == 0xFu, own_lo = own == 0x3u, own_hi = own == 0xCu;

  auto issue_box = [&](int box) {
    const int stage = box % NSTAGE;
    unsigned char* dst = ring + (size_t)stage * STAGE_BYTES;
    const int row0 = i_begin + box * RB;
    if (lane == 0) {
      mbar_expect_tx(&full_bar[warp][stage], (u_zero ? 0u : BOX64) + BOX64 + (has_e ? BOX32 : 0u));
      if (!u_zero) tma_load_2d(dst, &map_u, g0, row0, &full_bar[warp][stage]);
      tma_load_2d(dst + BOX64, &map_f, g0, row0, &full_bar[warp][stage]);
      if (has_e) tma_load_2d(dst + 2 * BOX64, &map_e, g0, row0, &full_bar[warp][stage]);
    }
  };
#pragma unroll
  for (int b = 0; b < NSTAGE; ++b)
    if (b < nbox) issue_box(b);

  // ---- state -----------------------------------------------------------------------------------
  double uw[3][4];   // u_new rows of ages 0, 1, 2
  double f64w[2][4]; // f rows of ages 0, 1
  float w[WR][4];    // e' window; age a = row (i - 1) - a
  float fr[FR][4];   // r32 window
  float rr[3][4];    // residual rows of the error equation (restriction)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) { uw[a][e] = 0.0; rr[a][e] = 0.f; }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) f64w[a][e] = 0.0;
#pragma unroll
  for (int a = 0; a < WR; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) w[a][e] = 0.f;
#pragma unroll
  for (int a = 0; a < FR; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) fr[a][e] = 0.f;
  double acc = 0.0;

  const int jc0 = (g0 >> 1) + 2 * lane;  // coarse column of element 0

  auto process_box = [&](auto masked_tag, const int ib, const uint32_t su, const uint32_t sfa, const uint32_t se) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    double* urow = p.u_out + (int64_t)ib * p.ld_uo + jbase;               // row i
    float* rrow = p.r_out + (int64_t)(ib - 1) * p.ld_r + jbase;          // row i - 1
    float* erow = p.e_out + (int64_t)(ib - 1 - NS) * p.ld_eo + jbase;    // row i - 1 - NS
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int i = ib + k;          // newest row of the fp64 part; parity of i == parity of k
      const int ip = i - 1;          // newest row of the fp32 pipeline
      const int kpar = (k + 1) & 1;  // parity of ip

      // (0) shift the windows
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uw[2][e] = uw[1][e]; uw[1][e] = uw[0][e];
        f64w[1][e] = f64w[0][e];
      }
#pragma unroll
      for (int a = WR - 1; a > 0; --a)
#pragma unroll
        for (int e = 0; e < 4; ++e) w[a][e] = w[a - 1][e];
#pragma unroll
      for (int a = FR - 1; a > 0; --a)
#pragma unroll
        for (int e = 0; e < 4; ++e) fr[a][e] = fr[a - 1][e];

      // (1) newest row: u_new = u + (double) e   (rows / columns outside the domain arrive as zeros)
      if (u_zero) {
#pragma unroll
        for (int e = 0; e < 4; ++e) uw[0][e] = 0.0;
      } else {
        lds4(su + (uint32_t)(k * STRIP * 8), uw[0]);
      }
      lds4(sfa + (uint32_t)(k * STRIP * 8), f64w[0]);
      if (has_e) {
        float ev[4];
        lds4(se + (uint32_t)(k * STRIP * 4), ev);
        uw[0][0] += (double)ev[0]; uw[0][1] += (double)ev[1]; uw[0][2] += (double)ev[2]; uw[0][3] += (double)ev[3];
        const bool row_ok = (i >= I0 && i < I1);
        double* dst = urow;
        if (!MASKED) {
          stg4_if(row_ok && own_all, dst, uw[0][0], uw[0][1], uw[0][2], uw[0][3]);
          stg2_if(row_ok && own_lo, dst, uw[0][0], uw[0][1]);
          stg2_if(row_ok && own_hi, dst + 2, uw[0][2], uw[0][3]);
        } else if (row_ok && own != 0u) {
          if (own == 0xFu) {
            stg4(dst, uw[0]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if ((own >> e) & 1u) dst[e] = uw[0][e];
          }
        }
      }
      urow += p.ld_uo;

      // (2) fp64 residual of row ip = i - 1 (its neighbours i - 2, i are in the window), norm, fp32 rounding
      {
        double r[4] = {0.0, 0.0, 0.0, 0.0};
        if (!MASKED || (ip >= 0 && ip < nx)) {
          if (!MASKED || (ip >= 1 && ip <= nx - 2)) {
            const double lfx = shfl_up1(uw[1][3]);
            const double rtx = shfl_dn1(uw[1][0]);
            const double r0 = residual_sel<SIMPLE, double>(sd, uw[1][0], uw[0][0], uw[2][0], uw[1][1], lfx, f64w[1][0]);
            const double r1 = residual_sel<SIMPLE, double>(sd, uw[1][1], uw[0][1], uw[2][1], uw[1][2], uw[1][0], f64w[1][1]);
            const double r2 = residual_sel<SIMPLE, double>(sd, uw[1][2], uw[0][2], uw[2][2], uw[1][3], uw[1][1], f64w[1][2]);
            const double r3 = residual_sel<SIMPLE, double>(sd, uw[1][3], uw[0][3], uw[2][3], rtx, uw[1][2], f64w[1][3]);
            r[0] = (!MASKED || (upd & 1u)) ? r0 : f64w[1][0];
            r[1] = (!MASKED || (upd & 2u)) ? r1 : f64w[1][1];
            r[2] = (!MASKED || (upd & 4u)) ? r2 : f64w[1][2];
            r[3] = (!MASKED || (upd & 8u)) ? r3 : f64w[1][3];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = f64w[1][e];  // boundary rows: r = f (laplacian.py:64,117)
          }
        }
        const bool row_ok = (ip >= I0 && ip < I1);
        if (row_ok && ip >= p.norm_row_lo && ip < p.norm_row_hi) {
          double rowsum;
          if (!MASKED && own == 0xFu) {
            rowsum = (r[0] * r[0] + r[1] * r[1]) + (r[2] * r[2] + r[3] * r[3]);
          } else {
            const double m0 = (own & 1u) ? r[0] : 0.0, m1 = (own & 2u) ? r[1] : 0.0;
            const double m2 = (own & 4u) ? r[2] : 0.0, m3 = (own & 8u) ? r[3] : 0.0;
            rowsum = (m0 * m0 + m1 * m1) + (m2 * m2 + m3 * m3);
          }
          acc += rowsum;
        }
        // (3a) the fp32 residual row: newest right-hand side row of the smoothing pipeline, and an output
        fr[0][0] = (float)r[0]; fr[0][1] = (float)r[1]; fr[0][2] = (float)r[2]; fr[0][3] = (float)r[3];
        float* dst = rrow;
        if (!MASKED) {
          stg4_if(row_ok && own_all, dst, fr[0][0], fr[0][1], fr[0][2], fr[0][3]);
          stg2_if(row_ok && own_lo, dst, fr[0][0], fr[0][1]);
          stg2_if(row_ok && own_hi, dst + 2, fr[0][2], fr[0][3]);
        } else if (row_ok && own != 0u) {
          if (own == 0xFu) {
            stg4(dst, fr[0]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if ((own >> e) & 1u) dst[e] = fr[0][e];
          }
        }
        rrow += p.ld_r;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) w[0][e] = 0.f;  // the error equation starts from zero

      // (3b) the NS half-sweep stages on the fp32 window, stage s on row ip - s
#pragma unroll
      for (int s = 1; s <= NS; ++s) {
        const int q = ip - s;
        if (!MASKED || (q >= 1 && q <= nx - 2)) {
          const int e0 = (kpar + s + ((s - 1) & 1)) & 1;
          if (e0 == 0) {
            const float lfx = shfl_up1(w[s][3]);
            const float n0 = relax_sel<SIMPLE, float>(sf, w[s][0], w[s - 1][0], w[s + 1][0], w[s][1], lfx, fr[s][0]);
            const float n2 = relax_sel<SIMPLE, float>(sf, w[s][2], w[s - 1][2], w[s + 1][2], w[s][3], w[s][1], fr[s][2]);
            w[s][0] = (!MASKED || (upd & 1u)) ? n0 : w[s][0];
            w[s][2] = (!MASKED || (upd & 4u)) ? n2 : w[s][2];
          } else {
            const float rtx = shfl_dn1(w[s][0]);
            const float n1 = relax_sel<SIMPLE, float>(sf, w[s][1], w[s - 1][1], w[s + 1][1], w[s][2], w[s][0], fr[s][1]);
            const float n3 = relax_sel<SIMPLE, float>(sf, w[s][3], w[s - 1][3], w[s + 1][3], rtx, w[s][2], fr[s][3]);
            w[s][1] = (!MASKED || (upd & 2u)) ? n1 : w[s][1];
            w[s][3] = (!MASKED || (upd & 8u)) ? n3 : w[s][3];
          }
        }
      }

      // (3c) row ip - NS of the error iterate is final
      {
        const int qf = ip - NS;
        const bool row_ok = (qf >= I0 && qf < I1);
        float* dst = erow;
        erow += p.ld_eo;
        if (!MASKED) {
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

      // (3d) residual of the error equation on row ip - NS - 1, full weighting every other row
      {
        constexpr int A = NS + 1;
        const int q2 = ip - A;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (!MASKED || (q2 >= 0 && q2 < nx)) {
          if (!MASKED || (q2 >= 1 && q2 <= nx - 2)) {
            const float lfx = shfl_up1(w[A][3]);
            const float rtx = shfl_dn1(w[A][0]);
            const float r0 = residual_sel<SIMPLE, float>(sf, w[A][0], w[A - 1][0], w[A + 1][0], w[A][1], lfx, fr[A][0]);
            const float r1 = residual_sel<SIMPLE, float>(sf, w[A][1], w[A - 1][1], w[A + 1][1], w[A][2], w[A][0], fr[A][1]);
            const float r2 = residual_sel<SIMPLE, float>(sf, w[A][2], w[A - 1][2], w[A + 1][2], w[A][3], w[A][1], fr[A][2]);
            const float r3 = residual_sel<SIMPLE, float>(sf, w[A][3], w[A - 1][3], w[A + 1][3], rtx, w[A][2], fr[A][3]);
            r[0] = (!MASKED || (upd & 1u)) ? r0 : fr[A][0];
            r[1] = (!MASKED || (upd & 2u)) ? r1 : fr[A][1];
            r[2] = (!MASKED || (upd & 4u)) ? r2 : fr[A][2];
            r[3] = (!MASKED || (upd & 8u)) ? r3 : fr[A][3];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = fr[A][e];
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          rr[2][e] = rr[1][e];
          rr[1][e] = rr[0][e];
          rr[0][e] = r[e];
        }
        if (kpar == (NS & 1)) {  // q2 is odd: rows q2-2, q2-1 (centre, even), q2 are complete
          const int fi = q2 - 1;
          const int ic = fi >> 1;
          if (fi >= I0 && fi < I1) {
            const float l2 = shfl_up1(rr[2][3]), l1 = shfl_up1(rr[1][3]), l0 = shfl_up1(rr[0][3]);
            const bool brow = MASKED && (ic == 0 || ic == p.nxc - 1);
            float v0, v1;
            {
              const float corners = ((l2 + rr[2][1]) + l0) + rr[0][1];
              const float edges = ((rr[2][0] + rr[0][0]) + l1) + rr[1][1];
              v0 = (0.0625f * corners + 0.125f * edges) + 0.25f * rr[1][0];
              const bool b = MASKED && (brow || jc0 == 0 || jc0 == p.nyc - 1);
              v0 = b ? rr[1][0] : v0;
            }
            {
              const float corners = ((rr[2][1] + rr[2][3]) + rr[0][1]) + rr[0][3];
              const float edges = ((rr[2][2] + rr[0][2]) + rr[1][1]) + rr[1][3];
              v1 = (0.0625f * corners + 0.125f * edges) + 0.25f * rr[1][2];
              const bool b = MASKED && (brow || jc0 + 1 == 0 || jc0 + 1 == p.nyc - 1);
              v1 = b ? rr[1][2] : v1;
            }
            const bool o0 = (own & 1u) != 0u, o1 = (own & 4u) != 0u;
            float* dst = p.coarse_out + (int64_t)ic * p.ld_co + jc0;
            stg2_if(o0 && o1, dst, v0, v1);
            stg1_if(o0 && !o1, dst, v0);
            stg1_if(o1 && !o0, dst + 1, v1);
          }
        }
      }
    }  // rows of the box
  };

  for (int box = 0; box < nbox; ++box) {
    const int stage = box % NSTAGE;
    mbar_wait(&full_bar[warp][stage], (uint32_t)((box / NSTAGE) & 1));
    const uint32_t sbox = ring_a + (uint32_t)stage * STAGE_BYTES;
    const uint32_t su = sbox + (uint32_t)(lane * LANE_V * 8);
    const uint32_t sfa = su + BOX64;
    const uint32_t se = sbox + 2 * BOX64 + (uint32_t)(lane * LANE_V * 4);
    const int ib = i_begin + box * RB;
    // interior fast path: the oldest row a box evaluates is the centre row of the restriction, ib - 1 - NS - 3
    constexpr int OLDEST = 1 + NS + 3;
    const bool fast = strip_interior && (ib - OLDEST >= 1) && (ib + RB - 1 <= nx - 2);
    if (fast) process_box(FalseTag{}, ib, su, sfa, se);
    else process_box(TrueTag{}, ib, su, sfa, se);
    __syncwarp();
    if (box + NSTAGE < nbox) issue_box(box + NSTAGE);
  }

  acc = warp_sum(acc);
  if (lane == 0) p.partials[(size_t)blockIdx.y * (gridDim.x * WARPS) + strip] = acc;
}

}  // namespace stream
}  // namespace mg
