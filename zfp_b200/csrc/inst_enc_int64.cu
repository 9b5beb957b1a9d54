// encode kernels for int64 (all dims, output modes, lossy + reversible)
#include "inst.cuh"
namespace zb {
template <> cudaError_t launch_encode_t<2>(int dims, int out_mode, const EncodeArgs& a) { return launch_encode_impl<2>(dims, out_mode, a); }
template <> cudaError_t launch_encode_var1_t<2>(const EncodeArgs& a, const Var1Bufs& v) { return launch_encode_var1_impl<2>(a, v); }
template <> int var1_tile_blocks<2>() { return EncCfg<2>::threads; }
}
