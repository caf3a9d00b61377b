// Instantiations + launcher of the fused defect + down pass (mg_stream_dd.cuh).
#include <stdlib.h>
#include "mg_stream_inst.cuh"
#include "mg_stream_dd.cuh"

namespace mg {
namespace stream {

constexpr int DD_WARPS = 4;

// Ring shapes (rows per box, boxes per warp).  MG_DD_VARIANT selects one at load time; it exists for tuning runs
// (tools/bench_dd.py), the default is the measured best.
struct DDVariant { int rb, nstage; };
static const DDVariant kVariants[] = {{2, 3}, {4, 2}};  // measured on 16385^2: 1.715 / 1.725 ms
static int dd_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MG_DD_VARIANT");
    v = e ? atoi(e) : 0;
    if (v < 0 || v >= (int)(sizeof(kVariants) / sizeof(kVariants[0]))) v = 0;
  }
  return v;
}
int defect_down_box_rows() { return kVariants[dd_variant()].rb; }

template <bool SIMPLE, int RB_, int NSTAGE_, int MINB_ = 1>
static int launch_dd(const CUtensorMap& mu, const CUtensorMap& mf, const CUtensorMap& me, DDParams& p,
                     const StencilScalars<double>& sd, const StencilScalars<float>& sf, cudaStream_t st) {
  auto kern = defect_down_kernel<SIMPLE, DD_WARPS, NSTAGE_, RB_, MINB_>;
  constexpr size_t stage_bytes = 2 * (size_t)RB_ * STRIP * 8 + (size_t)RB_ * STRIP * 4;
  constexpr size_t smem = (size_t)DD_WARPS * NSTAGE_ * stage_bytes;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[dev] = true;
  }
  if (p.rows_per_tile <= 0)
    p.rows_per_tile = pick_rows(p.nx, p.nstrips, DDGeometry::ROW_LEAD + DDGeometry::ROW_TAIL + 2);
  const int ntiles = (p.nx + p.rows_per_tile - 1) / p.rows_per_tile;
  dim3 grid((p.nstrips + DD_WARPS - 1) / DD_WARPS, ntiles);
  kern<<<grid, DD_WARPS * 32, smem, st>>>(mu, mf, me, p, sd, sf);
  return 0;
}

int launch_defect_down(bool simple, const CUtensorMap& mu, const CUtensorMap& mf, const CUtensorMap& me, DDParams& p,
                       const StencilScalars<double>& sd, const StencilScalars<float>& sf, cudaStream_t st) {
#define MG_DD_CASE(V, RB_, NS_, MB_)                                                                \
  if (dd_variant() == V)                                                                            \
    return simple ? launch_dd<true, RB_, NS_, MB_>(mu, mf, me, p, sd, sf, st)                       \
                  : launch_dd<false, RB_, NS_, MB_>(mu, mf, me, p, sd, sf, st);
  MG_DD_CASE(0, 2, 3, 1)
  MG_DD_CASE(1, 4, 2, 1)
#undef MG_DD_CASE
  return MG_ERR_BADARG;
}

}  // namespace stream
}  // namespace mg
