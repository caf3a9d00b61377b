// Instantiations of the streaming kernel for T = float with variable coefficients (TMA loader).
#include "mg_stream_inst.cuh"
namespace mg { namespace stream {
int launch_pass_f32_var(int nu, int front, int back, bool noblend, const Maps& m, PassParams& p,
                       const StencilScalars<float>& sc, cudaStream_t st) {
  return launch_pass_var<float>(nu, front, back, noblend, m, p, sc, st);
}
}}  // namespace mg::stream
