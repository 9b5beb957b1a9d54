// encode kernels for double (all dims, output modes, lossy + reversible)
#include "inst.cuh"
namespace zb {
template <> cudaError_t launch_encode_t<4>(int dims, int out_mode, const EncodeArgs& a) { return launch_encode_impl<4>(dims, out_mode, a); }
template <> cudaError_t launch_encode_var1_t<4>(const EncodeArgs& a, const Var1Bufs& v) { return launch_encode_var1_impl<4>(a, v); }
template <> int var1_tile_blocks<4>() { return EncCfg<4>::threads; }
}
