// bitstream_impl.h - in-memory layout of `struct bitstream`, shared by the host API and the backend.
//
// The layout is the reference's (include/zfp/bitstream.inl:133-143, default build: 64-bit words,
// no BIT_STREAM_STRIDED) because the reference's CUDA translation unit reads and pokes these
// fields directly (src/cuda_zfp/cuZFP.cu:20-26, 406-411) and a drop-in backend has to leave the
// struct in the state the reference's zfp_compress / zfp_decompress epilogue expects.
#pragma once

#include <cstddef>
#include <cstdint>

extern "C" {
struct bitstream {
  size_t bits;      // number of buffered bits, 0 <= bits < 64
  uint64_t buffer;  // buffered bits (buffer < 2^bits)
  uint64_t* ptr;    // next word to be read / written
  uint64_t* begin;  // first word of the stream
  uint64_t* end;    // one past the last word (capacity; not enforced)
};
}
