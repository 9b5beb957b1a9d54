// decode kernels for float (all dims, offset modes, lossy + reversible)
#include "inst.cuh"
namespace zb {
template <> cudaError_t launch_decode_t<3>(int dims, int offs_mode, const DecodeArgs& a) { return launch_decode_impl<3>(dims, offs_mode, a); }
template <> cudaError_t launch_index_t<3>(int dims, const DecodeArgs& a, uint16_t* lengths) { return launch_index_impl<3>(dims, a, lengths); }
template <> cudaError_t launch_spec_index_t<3>(int dims, int pass, const SpecIndexArgs& a, cudaStream_t st) { return launch_spec_index_impl<3>(dims, pass, a, st); }
}
