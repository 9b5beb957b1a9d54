// encode kernels for float (all dims, output modes, lossy + reversible)
#include "inst.cuh"
namespace zb {
template <> cudaError_t launch_encode_t<3>(int dims, int out_mode, const EncodeArgs& a) { return launch_encode_impl<3>(dims, out_mode, a); }
template <> cudaError_t launch_encode_var1_t<3>(const EncodeArgs& a, const Var1Bufs& v) { return launch_encode_var1_impl<3>(a, v); }
template <> int var1_tile_blocks<3>() { return EncCfg<3>::threads; }
}
