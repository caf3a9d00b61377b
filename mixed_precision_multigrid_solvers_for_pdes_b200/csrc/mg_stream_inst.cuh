// Host-side launcher for rbgs_stream_kernel, instantiated per (T, LOADER) translation unit to
// keep compile times parallel.  Included by mg_stream_f32_tma.cu etc. with MG_T / MG_LOADER set.
#pragma once
#include "mg_stream.cuh"

namespace mg {
namespace stream {

constexpr int WARPS = 4;
constexpr int RB = 4;

template <typename T> struct Stages { static constexpr int N = 4; };
template <> struct Stages<double> { static constexpr int N = 3; };

template <typename T, int NU, bool PROLONG, int BACK, int LOADER>
static int launch_one(const CUtensorMap& mu, const CUtensorMap& mf, const PassParams& p, const StencilScalars<T>& sc,
                      cudaStream_t st) {
  constexpr int NS = Stages<T>::N;
  auto kern = rbgs_stream_kernel<T, NU, PROLONG, BACK, LOADER, WARPS, NS, RB>;
  constexpr size_t smem = (size_t)WARPS * NS * 2 * RB * STRIP * sizeof(T);
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = true;
  }
  const int ntiles = (p.nx + p.rows_per_tile - 1) / p.rows_per_tile;
  dim3 grid((p.nstrips + WARPS - 1) / WARPS, ntiles);
  kern<<<grid, WARPS * 32, smem, st>>>(mu, mf, p, sc);
  return 0;
}

template <typename T, int LOADER>
int launch_pass(int nu, bool prolong, int back, const CUtensorMap& mu, const CUtensorMap& mf, const PassParams& p,
                const StencilScalars<T>& sc, cudaStream_t st) {
#define MG_CASE(NU_, PR_, BK_) \
  if (nu == NU_ && prolong == PR_ && back == BK_) return launch_one<T, NU_, PR_, BK_, LOADER>(mu, mf, p, sc, st);
  MG_CASE(0, false, BACK_RESTRICT) MG_CASE(0, false, BACK_NORM)
  MG_CASE(0, true, BACK_NONE) MG_CASE(0, true, BACK_RESTRICT) MG_CASE(0, true, BACK_NORM)
  MG_CASE(1, false, BACK_NONE) MG_CASE(1, false, BACK_RESTRICT) MG_CASE(1, false, BACK_NORM)
  MG_CASE(1, true, BACK_NONE) MG_CASE(1, true, BACK_RESTRICT) MG_CASE(1, true, BACK_NORM)
  MG_CASE(2, false, BACK_NONE) MG_CASE(2, false, BACK_RESTRICT) MG_CASE(2, false, BACK_NORM)
  MG_CASE(2, true, BACK_NONE) MG_CASE(2, true, BACK_RESTRICT) MG_CASE(2, true, BACK_NORM)
#undef MG_CASE
  return MG_ERR_UNSUPPORTED;
}

}  // namespace stream
}  // namespace mg
