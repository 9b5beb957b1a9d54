// encode kernels for int32 (all dims, output modes, lossy + reversible)
#include "inst.cuh"
namespace zb {
template <> cudaError_t launch_encode_t<1>(int dims, int out_mode, const EncodeArgs& a) { return launch_encode_impl<1>(dims, out_mode, a); }
template <> cudaError_t launch_encode_var1_t<1>(const EncodeArgs& a, const Var1Bufs& v) { return launch_encode_var1_impl<1>(a, v); }
template <> int var1_tile_blocks<1>() { return EncCfg<1>::threads; }
}
