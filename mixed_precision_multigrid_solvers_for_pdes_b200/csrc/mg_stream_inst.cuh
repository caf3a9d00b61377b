// Host-side launcher for rbgs_stream_kernel, instantiated per (T, LOADER) translation unit to
// keep compile times parallel.  Included by mg_stream_f32_tma.cu etc.
#pragma once
#include "mg_stream.cuh"

namespace mg {
namespace stream {

constexpr int WARPS = 4;
constexpr int RB = 4;

// ring depth: 3 boxes of 4 rows in flight per warp (fp32: 12 KB/warp), 2 for fp64 (16 KB/warp)
template <typename T> struct Stages { static constexpr int N = 3; };
template <> struct Stages<double> { static constexpr int N = 2; };

// Rows per tile.  A warp streams R + lead + tail rows sequentially, so a launch lasts about
// waves * (R + overlap) row-steps, with waves = warps needed / warps resident.  Large grids want tall tiles
// (overlap amortised), small grids want short ones (short critical path, all SMs busy) -- but never taller than
// 64 rows.  Measured (tools/bench_dd.py; profiles/r02_dd_tile_sweep.log, r02_rows_sweep_all_passes.log,
// r02_rows_sweep_coarse_levels.log): on 16385^2 every pass is flat between 48 and 96 rows, 1-2 % slower at 128, +3..12 %
// at 256 and +7..33 % at 512 rows (tiles that fill whole waves exactly, e.g. 514 rows, too); on 8193^2 (level 1 of
// the benchmarked cycle, level 0 of the heat runs) 128 rows cost 4-13 % against 48-64; on 4097^2 32-64 rows are
// best.  Short tiles keep the rows that are streamed concurrently close together (column halos hit L2) and leave
// a short tail when the last wave of CTAs is only partly filled.
static inline int pick_rows(int nx, int nstrips, int overlap) {
  const int64_t capacity = (int64_t)sm_count() * 12;  // resident warps (3 blocks of 4 warps per SM)
  int best = 64;
  int64_t best_cost = INT64_MAX;
  for (int r = 64; r >= 8; r >>= 1) {
    const int64_t warps = (int64_t)nstrips * ((nx + r - 1) / r);
    const int64_t waves = (warps + capacity - 1) / capacity;
    const int64_t cost = waves * (r + overlap);
    if (cost < best_cost) { best_cost = cost; best = r; }
  }
  return best;
}

struct Maps {
  CUtensorMap u, f, e, a;
};

template <typename T, int NU, int FRONT, int BACK, int LOADER, bool SIMPLE, int SMOOTH = SMOOTH_RBGS, bool VARCOEF = false>
static int launch_one(const Maps& m, PassParams& p, const StencilScalars<T>& sc, cudaStream_t st) {
  // fp32 prolongation passes carry the coarse slab in every stage: 2 stages keep 5 blocks (20 warps) per SM
  constexpr int NS = ((sizeof(T) == 4 && FRONT == FRONT_PROLONG && LOADER == LOADER_TMA) || VARCOEF) ? 2 : Stages<T>::N;
  auto kern = rbgs_stream_kernel<T, NU, FRONT, BACK, LOADER, SIMPLE, SMOOTH, WARPS, NS, RB, VARCOEF>;
  constexpr bool stage_coarse = StageCoarse<T, FRONT, BACK, LOADER>::value;
  constexpr size_t cbox = stage_coarse ? ((((size_t)(RB / 2 + 1) * COARSE_BOX_W * sizeof(T)) + 127) & ~(size_t)127) : 0;
  constexpr size_t stage_bytes = (VARCOEF ? 3 : 2) * (size_t)RB * STRIP * sizeof(T) +
                                 (FRONT == FRONT_ADDFINE ? (size_t)RB * STRIP * 4 : 0) + cbox;
  constexpr size_t smem = (size_t)WARPS * NS * stage_bytes;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = true;
  }
  if (p.rows_per_tile <= 0) p.rows_per_tile = pick_rows(p.nx, p.nstrips, p.tile_overlap);
  const int ntiles = (p.nx + p.rows_per_tile - 1) / p.rows_per_tile;
  dim3 grid((p.nstrips + WARPS - 1) / WARPS, ntiles);
  kern<<<grid, WARPS * 32, smem, st>>>(m.u, m.f, m.e, m.a, p, sc);
  return 0;
}

// smooth: SMOOTH_RBGS or SMOOTH_JACOBI (damped Jacobi sweeps; TMA loader only, and always the general point
// update: omega = 1 Jacobi is not a smoother, so it gets no specialised instantiation)
template <typename T, int LOADER>
int launch_pass(int nu, int front, int back, bool simple, int smooth, const Maps& m, PassParams& p,
                const StencilScalars<T>& sc, cudaStream_t st) {
  if (smooth == SMOOTH_JACOBI && nu > 0) {
    if constexpr (LOADER == LOADER_TMA) {
#define MG_JCASE(NU_, FR_, BK_) \
  if (nu == NU_ && front == FR_ && back == BK_) return launch_one<T, NU_, FR_, BK_, LOADER, false, SMOOTH_JACOBI>(m, p, sc, st);
      MG_JCASE(1, FRONT_NONE, BACK_NONE) MG_JCASE(1, FRONT_NONE, BACK_RESTRICT) MG_JCASE(1, FRONT_NONE, BACK_NORM)
      MG_JCASE(1, FRONT_PROLONG, BACK_NONE) MG_JCASE(1, FRONT_PROLONG, BACK_RESTRICT) MG_JCASE(1, FRONT_PROLONG, BACK_NORM)
      MG_JCASE(2, FRONT_NONE, BACK_NONE) MG_JCASE(2, FRONT_NONE, BACK_RESTRICT) MG_JCASE(2, FRONT_NONE, BACK_NORM)
      MG_JCASE(2, FRONT_PROLONG, BACK_NONE) MG_JCASE(2, FRONT_PROLONG, BACK_RESTRICT) MG_JCASE(2, FRONT_PROLONG, BACK_NORM)
#undef MG_JCASE
    }
    return MG_ERR_UNSUPPORTED;
  }
#define MG_CASE(NU_, FR_, BK_)                                                           \
  if (nu == NU_ && front == FR_ && back == BK_)                                          \
    return simple ? launch_one<T, NU_, FR_, BK_, LOADER, true>(m, p, sc, st)             \
                  : launch_one<T, NU_, FR_, BK_, LOADER, false>(m, p, sc, st);
  MG_CASE(0, FRONT_NONE, BACK_RESTRICT) MG_CASE(0, FRONT_NONE, BACK_NORM)
  MG_CASE(0, FRONT_PROLONG, BACK_NONE) MG_CASE(0, FRONT_PROLONG, BACK_RESTRICT) MG_CASE(0, FRONT_PROLONG, BACK_NORM)
  MG_CASE(1, FRONT_NONE, BACK_NONE) MG_CASE(1, FRONT_NONE, BACK_RESTRICT) MG_CASE(1, FRONT_NONE, BACK_NORM)
  MG_CASE(1, FRONT_PROLONG, BACK_NONE) MG_CASE(1, FRONT_PROLONG, BACK_RESTRICT) MG_CASE(1, FRONT_PROLONG, BACK_NORM)
  MG_CASE(2, FRONT_NONE, BACK_NONE) MG_CASE(2, FRONT_NONE, BACK_RESTRICT) MG_CASE(2, FRONT_NONE, BACK_NORM)
  MG_CASE(2, FRONT_PROLONG, BACK_NONE) MG_CASE(2, FRONT_PROLONG, BACK_RESTRICT) MG_CASE(2, FRONT_PROLONG, BACK_NORM)
  if constexpr (sizeof(T) == 8) {  // mixed-precision defect-correction passes (fp64 iterate, fp32 correction/residual)
    MG_CASE(0, FRONT_NONE, BACK_RESID) MG_CASE(0, FRONT_ADDFINE, BACK_RESID) MG_CASE(0, FRONT_ADDFINE, BACK_NONE)
  }
#undef MG_CASE
  return MG_ERR_UNSUPPORTED;
}

// Variable-coefficient passes (-div(a grad u) + shift*u): red-black GS, TMA-staged; `noblend` = (omega == 1).
// fp64 keeps to one sweep per pass: a 2-sweep fp64 pass would hold 188 registers of row windows (u, f and the
// coefficient rows with their shuffled neighbour columns) before any temporaries.
template <typename T>
int launch_pass_var(int nu, int front, int back, bool noblend, const Maps& m, PassParams& p,
                    const StencilScalars<T>& sc, cudaStream_t st) {
#define MG_VCASE(NU_, FR_, BK_)                                                                              \
  if (nu == NU_ && front == FR_ && back == BK_)                                                              \
    return noblend ? launch_one<T, NU_, FR_, BK_, LOADER_TMA, true, SMOOTH_RBGS, true>(m, p, sc, st)          \
                   : launch_one<T, NU_, FR_, BK_, LOADER_TMA, false, SMOOTH_RBGS, true>(m, p, sc, st);
  MG_VCASE(0, FRONT_NONE, BACK_RESTRICT) MG_VCASE(0, FRONT_NONE, BACK_NORM) MG_VCASE(0, FRONT_PROLONG, BACK_NONE)
  MG_VCASE(1, FRONT_NONE, BACK_NONE) MG_VCASE(1, FRONT_NONE, BACK_RESTRICT) MG_VCASE(1, FRONT_NONE, BACK_NORM)
  MG_VCASE(1, FRONT_PROLONG, BACK_NONE) MG_VCASE(1, FRONT_PROLONG, BACK_NORM)
  if constexpr (sizeof(T) == 4) {
    MG_VCASE(2, FRONT_NONE, BACK_NONE) MG_VCASE(2, FRONT_NONE, BACK_RESTRICT) MG_VCASE(2, FRONT_NONE, BACK_NORM)
    MG_VCASE(2, FRONT_PROLONG, BACK_NONE) MG_VCASE(2, FRONT_PROLONG, BACK_NORM)
  } else {  // mixed-precision defect-correction passes (fp64 iterate, fp32 correction / residual)
    MG_VCASE(0, FRONT_NONE, BACK_RESID) MG_VCASE(0, FRONT_ADDFINE, BACK_RESID) MG_VCASE(0, FRONT_ADDFINE, BACK_NONE)
  }
#undef MG_VCASE
  return MG_ERR_UNSUPPORTED;
}

}  // namespace stream
}  // namespace mg
