// Instantiations of the streaming kernel for T = float, loader = LOADER_CPASYNC.
#include "mg_stream_inst.cuh"
namespace mg { namespace stream {
int launch_pass_f32_cpa(int nu, int front, int back, bool simple, int smooth, const Maps& m, PassParams& p,
                       const StencilScalars<float>& sc, cudaStream_t st) {
  return launch_pass<float, LOADER_CPASYNC>(nu, front, back, simple, smooth, m, p, sc, st);
}
}}  // namespace mg::stream
