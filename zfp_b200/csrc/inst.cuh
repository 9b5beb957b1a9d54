// inst.cuh - launchers: runtime (dims, output mode, reversible) -> kernel template instance.
// Included by one translation unit per scalar type and direction (inst_<dir>_<type>.cu) so the
// 100+ kernel instances compile in parallel.
#pragma once

#include <cstdlib>

#include "kernels.cuh"
#include "kernels4d.cuh"
#include "kernels4q.cuh"
#include "kernels_ws.cuh"
#include "kernels_ps.cuh"
#include "kernels_q4.cuh"
#include "kernels_var1.cuh"

namespace zb {

constexpr uint32_t kStagedMaxBits = 4096;  // staging stays under ~17 KiB per warp

template <int TYPE, int DIMS>
constexpr size_t plane_smem_bytes()
{
  constexpr int N = 1 << (2 * DIMS);
  return (size_t)(kThreads / 32) * Traits<TYPE>::P * 32 * sizeof(typename PlaneWord<N>::type);
}

template <class K>
inline cudaError_t allow_smem(K kernel, size_t bytes)
{
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// the same, remembering per kernel instance AND per device what was already granted (the attribute is
// per device: a process that drives several GPUs must set it on each)
template <class K>
inline cudaError_t allow_smem_cached(K kernel, size_t bytes, size_t (&granted)[64])
{
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return allow_smem(kernel, bytes);
  if (bytes > granted[dev]) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    granted[dev] = bytes;
  }
  return cudaSuccess;
}


template <int TYPE> cudaError_t run_encode_q4(const EncodeArgs& a);

// developer knob (occupancy experiments): extra dynamic shared memory per CTA of the staged kernels, bytes
inline size_t smem_pad()
{
  static const size_t pad = getenv("ZFP_B200_SMEM_PAD") ? (size_t)atol(getenv("ZFP_B200_SMEM_PAD")) : 0;
  return pad;
}

template <int TYPE, int DIMS, bool REV>
cudaError_t run_encode_staged(const EncodeArgs& a)
{
  constexpr int N = 1 << (2 * DIMS);
  if constexpr (Traits<TYPE>::P == 64 && DIMS == 3 && !REV) {
    static const bool q4 = getenv("ZFP_B200_Q4") != nullptr;
    if (q4 && (a.prm.maxbits & 63) == 0 && (a.start_bit & 63) == 0)
      return run_encode_q4<TYPE>(a);
  }
  auto kernel = encode_staged_kernel<TYPE, DIMS, REV>;
  constexpr int threads = SEncCfg<TYPE, DIMS>::threads;
  const size_t smem = (size_t)(threads / 32) * (kStagedPlanes * 32 * sizeof(typename PlaneWord<N>::type) +
                                                ((a.prm.maxbits >> 5) + kStageSlack) * 32 * 4) +
                      (kEncSmall8<TYPE, N, REV> ? kEncLut8Words * 4 : 0) +
                      (N == 4 ? kEncLut4Words * 4 : 0) + smem_pad();
  static size_t granted[64] = { 0 };  // per kernel instance
  cudaError_t e = allow_smem_cached(kernel, smem, granted);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.g.nblocks + threads - 1) / threads;
  kernel<<<(unsigned)ctas, threads, smem, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                    static_cast<uint64_t*>(a.out), a.start_bit);
  return cudaGetLastError();
}

template <int TYPE, int DIMS>
constexpr size_t var_smem_bytes()
{
  constexpr int N = 1 << (2 * DIMS);
  return (size_t)(kThreads / 32) * (kStagedPlanes * 32 * sizeof(typename PlaneWord<N>::type) + kVarStageWords * 32 * 4);
}

template <int TYPE, int DIMS, bool REV>
cudaError_t run_encode_var(const EncodeArgs& a)
{
  auto kernel = encode_var_kernel<TYPE, DIMS, REV>;
  constexpr int threads = EncCfg<TYPE>::threads;
  constexpr size_t smem = var_smem_bytes<TYPE, DIMS>() / (kThreads / 32) * (threads / 32);
  cudaError_t e = allow_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + threads - 1) / threads;
  kernel<<<(unsigned)ctas, threads, smem, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                    static_cast<uint32_t*>(a.out), a.slot_words * 2, a.lengths, a.b0, a.b1);
  return cudaGetLastError();
}

template <int TYPE, int DIMS, bool REV>
cudaError_t run_decode_var(const DecodeArgs& a)
{
  auto kernel = decode_var_kernel<TYPE, DIMS, REV>;
  constexpr int threads = DecCfg<TYPE>::threads;
  constexpr size_t smem = var_smem_bytes<TYPE, DIMS>() / (kThreads / 32) * (threads / 32);
  cudaError_t e = allow_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + threads - 1) / threads;
  kernel<<<(unsigned)ctas, threads, smem, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                    static_cast<const uint32_t*>(a.in), a.offsets, a.lengths, a.b0, a.b1, a.check);
  return cudaGetLastError();
}

template <int TYPE, int DIMS, int OUT, bool REV>
cudaError_t run_encode(const EncodeArgs& a)
{
  if constexpr (OUT == 0)
    if (a.staged && a.prm.maxbits <= kStagedMaxBits && a.b0 == 0 && a.b1 == a.g.nblocks)
      return run_encode_staged<TYPE, DIMS, REV>(a);
  // variable rate, reversible included (with the empty-plane shortcut the lockstep kernels match or
  // beat the general kernel's per-plane run loop for every type; measured at 1024^3 / 512^3)
  if constexpr (OUT == 2)
    if (a.staged)
      return run_encode_var<TYPE, DIMS, REV>(a);
  auto kernel = encode_kernel<TYPE, DIMS, OUT, REV>;
  constexpr size_t smem = plane_smem_bytes<TYPE, DIMS>();
  cudaError_t e = allow_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + kThreads - 1) / kThreads;
  kernel<<<(unsigned)ctas, kThreads, smem, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm, a.out,
                                                    a.start_bit, a.slot_words, a.lengths, a.b0, a.b1);
  return cudaGetLastError();
}

// multiprocessors of the current device
inline int device_sms()
{
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;
  return n;
}

// warp-specialised decoder (kernels_ws.cuh): persistent CTAs, two per multiprocessor
template <int TYPE>
cudaError_t run_decode_ws(const DecodeArgs& a)
{
  auto kernel = decode_ws_kernel<TYPE>;
  const size_t smem = ws_cta_bytes(a.prm.maxbits >> 5);
  static size_t granted[64] = { 0 };
  cudaError_t e = allow_smem_cached(kernel, smem, granted);
  if (e != cudaSuccess) return e;
  const uint64_t nbatches = (a.b1 - a.b0 + 31) / 32;
  uint64_t ctas = (nbatches + kWsPairs - 1) / kWsPairs;
  const uint64_t resident = (uint64_t)device_sms() * 2;
  if (ctas > resident) ctas = resident;
  kernel<<<(unsigned)ctas, kWsThreads, smem, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                      static_cast<const uint64_t*>(a.in), a.start_bit, a.b0, a.b1);
  return cudaGetLastError();
}

// phased register budget (kernels_ps.cuh): persistent CTAs, one per multiprocessor
template <int TYPE>
cudaError_t run_decode_ps(const DecodeArgs& a)
{
  auto kernel = decode_ps_kernel<TYPE>;
  const size_t smem = ps_cta_bytes(a.prm.maxbits >> 5);
  static size_t granted[64] = { 0 };
  cudaError_t e = allow_smem_cached(kernel, smem, granted);
  if (e != cudaSuccess) return e;
  const uint64_t nbatches = (a.b1 - a.b0 + kPsGroupThreads - 1) / kPsGroupThreads;
  uint64_t ctas = (nbatches + kPsGroups - 1) / kPsGroups;
  const uint64_t resident = (uint64_t)device_sms();
  if (ctas > resident) ctas = resident;
  kernel<<<(unsigned)ctas, kPsThreads, smem, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                      static_cast<const uint64_t*>(a.in), a.start_bit, a.b0, a.b1);
  return cudaGetLastError();
}

// four-lanes-per-block encoder (kernels_q4.cuh)
template <int TYPE>
cudaError_t run_encode_q4(const EncodeArgs& a)
{
  auto kernel = encode_q4_kernel<TYPE>;
  const size_t smem = (size_t)(kQ4Threads / 32) * q4_warp_bytes(a.prm.maxbits >> 5);
  static size_t granted[64] = { 0 };
  cudaError_t e = allow_smem_cached(kernel, smem, granted);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.g.nblocks + kQ4Threads - 1) / kQ4Threads;
  kernel<<<(unsigned)ctas, kQ4Threads, smem, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                      static_cast<uint64_t*>(a.out), a.start_bit);
  return cudaGetLastError();
}

// four-lanes-per-block decoder (kernels_q4.cuh)
template <int TYPE>
cudaError_t run_decode_q4(const DecodeArgs& a)
{
  auto kernel = decode_q4_kernel<TYPE>;
  const size_t smem = (size_t)(kQ4Threads / 32) * q4_warp_bytes(a.prm.maxbits >> 5);
  static size_t granted[64] = { 0 };
  cudaError_t e = allow_smem_cached(kernel, smem, granted);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + kQ4Threads - 1) / kQ4Threads;
  kernel<<<(unsigned)ctas, kQ4Threads, smem, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                      static_cast<const uint64_t*>(a.in), a.start_bit, a.b0, a.b1);
  return cudaGetLastError();
}

template <int TYPE, int DIMS, bool REV>
cudaError_t run_decode_staged(const DecodeArgs& a)
{
  constexpr int N = 1 << (2 * DIMS);
  if constexpr (Traits<TYPE>::P == 64 && DIMS == 3 && !REV) {
    static const bool q4 = getenv("ZFP_B200_Q4") != nullptr;
    if (q4 && !a.g.box && (kQ4Threads / 32) * q4_warp_bytes(a.prm.maxbits >> 5) <= 200 * 1024)
      return run_decode_q4<TYPE>(a);
    static const bool ps = getenv("ZFP_B200_PS") != nullptr;
    const uint32_t words = a.prm.maxbits >> 5;
    if (ps && (words & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.in) + (a.start_bit >> 6) * 8) & 15) == 0 &&
        ps_cta_bytes(words) <= 220 * 1024)
      return run_decode_ps<TYPE>(a);
    static const bool ws = getenv("ZFP_B200_WS") != nullptr && ws_cta_bytes(4096 >> 5) > 0;
    if (ws && !a.g.box && ws_cta_bytes(a.prm.maxbits >> 5) <= 110 * 1024)
      return run_decode_ws<TYPE>(a);
  }
  auto kernel = decode_staged_kernel<TYPE, DIMS, REV>;
  const size_t smem = (size_t)(SDecCfg<TYPE, DIMS>::threads / 32) * (kStagedPlanes * 32 * sizeof(typename PlaneWord<N>::type) +
                                                 ((a.prm.maxbits >> 5) + kReadSlack) * 32 * 4) +
                      (kDecSmall8<N, REV> ? kDecLut8hBytes : 0) + smem_pad();
  static size_t granted[64] = { 0 };  // per kernel instance
  cudaError_t e = allow_smem_cached(kernel, smem, granted);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + SDecCfg<TYPE, DIMS>::threads - 1) / SDecCfg<TYPE, DIMS>::threads;
  kernel<<<(unsigned)ctas, SDecCfg<TYPE, DIMS>::threads, smem, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                    static_cast<const uint64_t*>(a.in), a.start_bit, a.b0, a.b1);
  return cudaGetLastError();
}

template <int TYPE, int DIMS, int OFFS, bool REV>
cudaError_t run_decode(const DecodeArgs& a)
{
  if constexpr (OFFS == 0)
    // (blocks of an odd number of 32-bit words would work here too - the encoder takes them - but the general
    // kernel decodes them faster: 1-D fp64 at 8 bits/value 2.4 ms against 3.1 ms)
    if (a.staged && a.prm.maxbits <= kStagedMaxBits && (a.prm.maxbits & 63) == 0 && (a.start_bit & 63) == 0)
      return run_decode_staged<TYPE, DIMS, REV>(a);
  if constexpr (OFFS == 1)
    if (a.staged && a.lengths)
      return run_decode_var<TYPE, DIMS, REV>(a);
  auto kernel = decode_kernel<TYPE, DIMS, OFFS, REV>;
  constexpr size_t smem = DIMS == 1 ? (size_t)kDecLut4Bytes : plane_smem_bytes<TYPE, DIMS>();
  cudaError_t e = allow_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + kThreads - 1) / kThreads;
  kernel<<<(unsigned)ctas, kThreads, smem, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm, a.in,
                                                    a.start_bit, a.offsets, a.b0, a.b1, a.lengths, a.check);
  return cudaGetLastError();
}

template <int TYPE, int OUT>
cudaError_t run_encode4(const EncodeArgs& a)
{
  static const bool one_thread_per_block = getenv("ZFP_B200_4D_OLD") != nullptr;  // developer A/B switch
  if (!one_thread_per_block) {
    const unsigned ctas = (unsigned)((a.b1 - a.b0 + kBlocks4q - 1) / kBlocks4q);
    const size_t smem = (size_t)kBlocks4q * block_bytes4q<TYPE>();
    auto data = static_cast<const typename Traits<TYPE>::Scalar*>(a.data);
    if (a.prm.minexp < kMinExp)
      encode4q_kernel<TYPE, OUT, true><<<ctas, kThreads4q, smem, a.st>>>(data, a.g, a.prm, a.out, a.start_bit, a.slot_words, a.lengths, a.b0, a.b1);
    else
      encode4q_kernel<TYPE, OUT, false><<<ctas, kThreads4q, smem, a.st>>>(data, a.g, a.prm, a.out, a.start_bit, a.slot_words, a.lengths, a.b0, a.b1);
    return cudaGetLastError();
  }
  const unsigned ctas = (unsigned)((a.b1 - a.b0 + kThreads4 - 1) / kThreads4);
  encode4_kernel<TYPE, OUT><<<ctas, kThreads4, 0, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                           a.out, a.start_bit, a.slot_words, a.lengths, a.b0, a.b1);
  return cudaGetLastError();
}

template <int TYPE, int OFFS>
cudaError_t run_decode4(const DecodeArgs& a)
{
  static const bool one_thread_per_block = getenv("ZFP_B200_4D_OLD") != nullptr;
  if (!one_thread_per_block) {
    const unsigned ctas = (unsigned)((a.b1 - a.b0 + kBlocks4q - 1) / kBlocks4q);
    const size_t smem = (size_t)kBlocks4q * block_bytes4q<TYPE>();
    auto data = static_cast<typename Traits<TYPE>::Scalar*>(a.data);
    if (a.prm.minexp < kMinExp)
      decode4q_kernel<TYPE, OFFS, true><<<ctas, kThreads4q, smem, a.st>>>(data, a.g, a.prm, a.in, a.start_bit, a.offsets, a.b0, a.b1, a.lengths, a.check);
    else
      decode4q_kernel<TYPE, OFFS, false><<<ctas, kThreads4q, smem, a.st>>>(data, a.g, a.prm, a.in, a.start_bit, a.offsets, a.b0, a.b1, a.lengths, a.check);
    return cudaGetLastError();
  }
  const unsigned ctas = (unsigned)((a.b1 - a.b0 + kThreads4 - 1) / kThreads4);
  decode4_kernel<TYPE, OFFS><<<ctas, kThreads4, 0, a.st>>>(static_cast<typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm, a.in,
                                                            a.start_bit, a.offsets, a.b0, a.b1, a.lengths, a.check);
  return cudaGetLastError();
}

#define ZB_ENC_CASE(D, O)                                                     \
  case (D) * 10 + (O): return rev ? run_encode<TYPE, D, O, true>(a) : run_encode<TYPE, D, O, false>(a);
#define ZB_DEC_CASE(D, O)                                                     \
  case (D) * 10 + (O): return rev ? run_decode<TYPE, D, O, true>(a) : run_decode<TYPE, D, O, false>(a);

template <int TYPE>
cudaError_t launch_encode_impl(int dims, int out_mode, const EncodeArgs& a)
{
  const bool rev = a.prm.minexp < kMinExp;
  switch (dims * 10 + out_mode) {
    ZB_ENC_CASE(1, 0) ZB_ENC_CASE(1, 1) ZB_ENC_CASE(1, 2)
    ZB_ENC_CASE(2, 0) ZB_ENC_CASE(2, 1) ZB_ENC_CASE(2, 2)
    ZB_ENC_CASE(3, 0) ZB_ENC_CASE(3, 1) ZB_ENC_CASE(3, 2)
    case 40: return run_encode4<TYPE, 0>(a);
    case 41: return run_encode4<TYPE, 1>(a);
    case 42: return run_encode4<TYPE, 2>(a);
    default: return cudaErrorInvalidValue;
  }
}

template <int TYPE>
cudaError_t launch_decode_impl(int dims, int offs_mode, const DecodeArgs& a)
{
  const bool rev = a.prm.minexp < kMinExp;
  switch (dims * 10 + offs_mode) {
    ZB_DEC_CASE(1, 0) ZB_DEC_CASE(1, 1)
    ZB_DEC_CASE(2, 0) ZB_DEC_CASE(2, 1)
    ZB_DEC_CASE(3, 0) ZB_DEC_CASE(3, 1)
    case 40: return run_decode4<TYPE, 0>(a);
    case 41: return run_decode4<TYPE, 1>(a);
    default: return cudaErrorInvalidValue;
  }
}

template <int TYPE, bool REV>
cudaError_t run_encode_var1(const EncodeArgs& a, const Var1Bufs& v)
{
  if (v.cleanup) {
    auto again = reencode_kernel<TYPE, REV>;
    constexpr size_t smem2 = plane_smem_bytes<TYPE, 3>();
    cudaError_t e = allow_smem(again, smem2);
    if (e != cudaSuccess) return e;
    again<<<(unsigned)(v.sms * 4), kThreads, smem2, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm, a.out,
                                                             static_cast<const Var1Overflow*>(v.overflow), v.overflow_count, v.overflow_capacity);
    return cudaGetLastError();
  }
  auto kernel = encode_var1_kernel<TYPE, REV>;
  constexpr int threads = EncCfg<TYPE>::threads;
  constexpr size_t smem = var_smem_bytes<TYPE, 3>() / (kThreads / 32) * (threads / 32);
  cudaError_t e = allow_smem(kernel, smem);
  if (e != cudaSuccess) return e;
  const uint64_t ctas = (a.b1 - a.b0 + threads - 1) / threads;
  kernel<<<(unsigned)ctas, threads, smem, a.st>>>(static_cast<const typename Traits<TYPE>::Scalar*>(a.data), a.g, a.prm,
                                                    static_cast<uint64_t*>(a.out), a.lengths, a.b0, a.b1,
                                                    static_cast<Var1Status*>(v.status), v.ticket, v.carry,
                                                    static_cast<Var1Overflow*>(v.overflow), v.overflow_count, v.overflow_capacity);
  return cudaGetLastError();
}

template <int TYPE>
cudaError_t launch_encode_var1_impl(const EncodeArgs& a, const Var1Bufs& v)
{
  return a.prm.minexp < kMinExp ? run_encode_var1<TYPE, true>(a, v) : run_encode_var1<TYPE, false>(a, v);
}

template <int TYPE>
cudaError_t launch_index_impl(int dims, const DecodeArgs& a, uint16_t* lengths)
{
  const bool rev = a.prm.minexp < kMinExp;
  switch (dims) {
    case 1: if (rev) index_scan_kernel<TYPE, 1, true><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.b0, a.g.nblocks, a.prm, lengths);
            else index_scan_kernel<TYPE, 1, false><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.b0, a.g.nblocks, a.prm, lengths); break;
    case 2: if (rev) index_scan_kernel<TYPE, 2, true><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.b0, a.g.nblocks, a.prm, lengths);
            else index_scan_kernel<TYPE, 2, false><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.b0, a.g.nblocks, a.prm, lengths); break;
    case 3: if (rev) index_scan_kernel<TYPE, 3, true><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.b0, a.g.nblocks, a.prm, lengths);
            else index_scan_kernel<TYPE, 3, false><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.b0, a.g.nblocks, a.prm, lengths); break;
    case 4: if (a.b0) return cudaErrorInvalidValue;
            index_scan4_kernel<TYPE><<<1, 1, 0, a.st>>>(a.in, a.start_bit, a.g.nblocks, a.prm, lengths); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <int TYPE, int DIMS, bool REV>
cudaError_t launch_spec_index_pass(int pass, const SpecIndexArgs& a, cudaStream_t st)
{
  const unsigned ctas = (a.nseg + 1) / 2;  // two walkers (warps) per CTA
  if (pass == 0) spec_index_kernel<TYPE, DIMS, REV, 0><<<ctas, 64, 0, st>>>(a);
  else if (pass == 1) spec_index_kernel<TYPE, DIMS, REV, 1><<<ctas, 64, 0, st>>>(a);
  else spec_index_kernel<TYPE, DIMS, REV, 2><<<ctas, 64, 0, st>>>(a);
  return cudaGetLastError();
}

template <int TYPE>
cudaError_t launch_spec_index_impl(int dims, int pass, const SpecIndexArgs& a, cudaStream_t st)
{
  const bool rev = a.prm.minexp < kMinExp;
  switch (dims) {
    case 1: return rev ? launch_spec_index_pass<TYPE, 1, true>(pass, a, st) : launch_spec_index_pass<TYPE, 1, false>(pass, a, st);
    case 2: return rev ? launch_spec_index_pass<TYPE, 2, true>(pass, a, st) : launch_spec_index_pass<TYPE, 2, false>(pass, a, st);
    case 3: return rev ? launch_spec_index_pass<TYPE, 3, true>(pass, a, st) : launch_spec_index_pass<TYPE, 3, false>(pass, a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace zb
