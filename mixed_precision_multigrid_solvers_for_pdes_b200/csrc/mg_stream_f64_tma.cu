// Instantiations of the streaming kernel for T = double, loader = LOADER_TMA.
#include "mg_stream_inst.cuh"
namespace mg { namespace stream {
int launch_pass_f64_tma(int nu, int front, int back, bool simple, int smooth, const Maps& m, PassParams& p,
                       const StencilScalars<double>& sc, cudaStream_t st) {
  return launch_pass<double, LOADER_TMA>(nu, front, back, simple, smooth, m, p, sc, st);
}
}}  // namespace mg::stream
