// Instantiations of the streaming kernel for T = double, loader = LOADER_TMA.
#include "mg_stream_inst.cuh"
namespace mg { namespace stream {
int launch_pass_f64_tma(int nu, bool prolong, int back, const CUtensorMap& mu, const CUtensorMap& mf,
                       const PassParams& p, const StencilScalars<double>& sc, cudaStream_t st) {
  return launch_pass<double, LOADER_TMA>(nu, prolong, back, mu, mf, p, sc, st);
}
}}  // namespace mg::stream
