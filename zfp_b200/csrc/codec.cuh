// codec.cuh - per-block zfp codec for sm_100a, one 4^d block per thread (d <= 3).
//
// Work split inside a warp (32 blocks at a time):
//   * everything that is data parallel inside a block - strided gather, block-floating-point
//     cast, lifting transform, sequency reorder, negabinary - runs on the block's 4^d values held
//     in registers with compile-time indices (no local memory);
//   * bit planes are formed by an in-register 32x32 bit-matrix transpose (5 butterfly stages),
//     not by per-bit loops, and parked in shared memory as [plane][lane] so that every lane's
//     accesses are conflict free;
//   * the embedded coder walks the planes with the whole warp in lockstep ("plane strings": n
//     verbatim bits, then the group-tested part T of the plane); per plane a lane builds / parses T
//     with bit-parallel arithmetic instead of a loop over its items (encode_plane_lockstep,
//     decode_plane_lockstep), falling back to an exact per-item loop where that cannot be proven
//     right.  The general kernels (odd rates, stream offsets) use the plain per-run loops
//     (encode_planes / decode_planes).
//
// Reference semantics reproduced (paths in the reference tree):
//   gather/pad      src/template/encode{1,2,3}.c, encode.c:8-27
//   exponent/cast   src/template/encodef.c:10-59, codecf.c:5-32
//   lifting         src/template/encode.c:30-56, decode.c:8-45, revencode.c:6-38, revdecode.c:6-38
//   order/negabin   src/template/encode.c:75-88, decode.c:63-76, codec{1,2,3}.c
//   embedded coder  src/template/encode.c:91-256, decode.c:79-278
//   reversible      src/template/revencode.c:41-79, revencodef.c:6-80, revdecode*.c, revcodecf.c
#pragma once

#include <cstdint>
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>

#include "types.h"
#include "zfp_perm_tables.h"
#include "coder_luts.h"

namespace zb {

// 64-bit 3-D decode: 1 = the CTA's warps rendezvous before the straight-line tail (shared instruction fetch).  It paid
// while the tail was 1900 instructions of integer lifting; with the FP64-pipe tail (half as long) it does not any more
// (1024^3 fp64 rate 4 / 8 / 16 decompress with: 3.90 / 4.94 / 6.00 ms, without: 3.73 / 4.94 / 5.92 ms)
#ifndef ZB_DEC_TAIL_SYNC
#define ZB_DEC_TAIL_SYNC 0
#endif
#ifndef ZB_FP64_TAIL
#define ZB_FP64_TAIL 1  // inverse transform of fp64 blocks on the FP64 pipe where that is exact (decode_block)
#endif
// developer switch for A/B timing: 0 removes the narrow (32-bit) plane steps of the 64-value coders
#ifndef ZB_NARROW
#define ZB_NARROW 0
#endif

template <int TYPE> struct Traits;
template <> struct Traits<T_INT32> {
  using Scalar = int32_t; using Int = int32_t; using UInt = uint32_t;
  static constexpr int P = 32, EBITS = 0, EBIAS = 0, MANT = 0, PBITS = 5;
  static constexpr bool is_fp = false;
};
template <> struct Traits<T_INT64> {
  using Scalar = int64_t; using Int = int64_t; using UInt = uint64_t;
  static constexpr int P = 64, EBITS = 0, EBIAS = 0, MANT = 0, PBITS = 6;
  static constexpr bool is_fp = false;
};
template <> struct Traits<T_FLOAT> {
  using Scalar = float; using Int = int32_t; using UInt = uint32_t;
  static constexpr int P = 32, EBITS = 8, EBIAS = 127, MANT = 23, PBITS = 5;
  static constexpr bool is_fp = true;
};
template <> struct Traits<T_DOUBLE> {
  using Scalar = double; using Int = int64_t; using UInt = uint64_t;
  static constexpr int P = 64, EBITS = 11, EBIAS = 1023, MANT = 52, PBITS = 6;
  static constexpr bool is_fp = true;
};

__constant__ uint8_t c_perm4[256] = ZFP_B200_PERM4_INIT;  // 4-D path only (kernels4d.cuh)

// compile-time lookup (switch, no table in memory) so that reordering is pure register renaming
// once the loops are unrolled
template <int DIMS>
__host__ __device__ constexpr int perm_at(int i)
{
  if (DIMS == 1) { switch (i) { ZFP_B200_PERM1_CASES } }
  else if (DIMS == 2) { switch (i) { ZFP_B200_PERM2_CASES } }
  else { switch (i) { ZFP_B200_PERM3_CASES } }
}

// ------------------------------------------------------------------------------------------------
// small bit helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t lowmask64(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1); }
// count trailing zeros of a non-zero 64-bit value with 32-bit operations
__device__ __forceinline__ uint32_t ctz64(uint64_t x)
{
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  const uint32_t w = lo ? lo : hi;
  return (uint32_t)__ffs((int)w) - 1 + (lo ? 0u : 32u);
}

// ------------------------------------------------------------------------------------------------
// bit writer: LSB-first into 64-bit words (include/zfp/bitstream.inl:288-313 semantics).
// MODE 0: the destination words are private to this thread from the first bit on (word-aligned
//         start): plain stores.
// MODE 1: neighbouring blocks share words: destination pre-zeroed, every word is OR-merged with a
//         fire-and-forget reduction.
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct BitWriter {
  static constexpr bool kStaged = false;
  unsigned long long* w;
  uint64_t acc;
  uint32_t fill;

  __device__ __forceinline__ void init(void* words, uint64_t bitpos)
  {
    w = reinterpret_cast<unsigned long long*>(words) + (bitpos >> 6);
    fill = (uint32_t)(bitpos & 63);
    acc = 0;
  }
  __device__ __forceinline__ void emit(uint64_t word)
  {
    if (MODE == 0)
      *w = word;
    else if (word)
      atomicOr(w, (unsigned long long)word);
    w++;
  }
  // append the low len bits of v (v < 2^len, 0 <= len <= 64)
  __device__ __forceinline__ void put(uint64_t v, uint32_t len)
  {
    acc |= v << fill;
    uint32_t total = fill + len;
    if (total >= 64) {
      emit(acc);
      acc = fill ? v >> (64 - fill) : 0;
      total -= 64;
    }
    fill = total;
  }
  __device__ __forceinline__ void pad(uint32_t nbits)
  {
    while (nbits >= 64) { put(0, 64); nbits -= 64; }
    put(0, nbits);
  }
  __device__ __forceinline__ void flush()
  {
    if (fill) { emit(acc); acc = 0; fill = 0; }
  }
};

__device__ __forceinline__ uint32_t mask32(uint32_t len)  // low `len` bits set, 0 <= len <= 32 (larger values clamp)
{
  uint32_t r;
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r) : "r"(len));
  return r;
}

// PTX shifts clamp the shift amount at the register width (a shift by >= width gives 0), which is
// exactly what the coder wants at n == N; C++ shifts leave that case undefined.
__device__ __forceinline__ uint64_t shr64c(uint64_t x, uint32_t n) { uint64_t r; asm("shr.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
__device__ __forceinline__ uint64_t shl64c(uint64_t x, uint32_t n) { uint64_t r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
__device__ __forceinline__ uint32_t shr32c(uint32_t x, uint32_t n) { uint32_t r; asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }
__device__ __forceinline__ uint32_t shl32c(uint32_t x, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }

// Column writer (plane-lockstep coder): the block's bits are assembled as 32-bit words in a
// lane-private shared-memory column ([word][lane], bank = lane: conflict free whatever word each
// lane is at), addressed by bit position.  An append ORs the value into the partial word kept in
// `acc` and stores every word the value can touch UNCONDITIONALLY (a later append rewrites the
// partial word with more bits in it): no predicates, no flush test, and the word holding `bp` is
// always current in shared memory.  Words beyond it hold zeros or were never written; finish()
// zero-fills up to the block size and the kernel copies exactly that many words (truncation).
struct ColWriter {
  static constexpr bool kStaged = true;
  static constexpr bool kLockstep = true;
  uint32_t base;  // shared-space byte address of word 0 of this lane's column (words are 128 bytes apart)
  uint32_t acc;   // bits of the word containing bp that lie below bp
  uint32_t bp;    // bits appended so far (variable rate: since the last drain)
  // variable rate: the column is a window of `cap` words; when it runs low its complete words
  // leave for the block's slot in global memory (gdst) and `drained` counts their bits
  uint32_t* gdst;
  uint32_t drained, cap;
  // single-pass variable rate (kernels_var1.cuh): no slot - a block that outgrows the window just counts
  // its bits (`drained` > 0 afterwards marks it) and is encoded again, in place, by the clean-up kernel
  bool sink;
  // shared-space address of a copy of kEncLut8 (0: none): the small-universe plane steps, encode_planes_small8
  uint32_t lut;

  __device__ __forceinline__ void init(uint32_t* column, uint32_t* slot = nullptr, uint32_t capacity_words = 0, bool count_only = false)
  {
    lut = 0;
    base = (uint32_t)__cvta_generic_to_shared(column);
    acc = 0;
    bp = 0;
    gdst = count_only ? reinterpret_cast<uint32_t*>(8) : slot;  // (non-null: the window logic is on; never dereferenced)
    drained = 0;
    cap = capacity_words;
    sink = count_only;
  }
  // call at plane boundaries: a plane appends < 200 bits and an append stores two words ahead
  __device__ __forceinline__ void drain_if_low()
  {
    if (gdst && bp > (cap - 10) * 32)
      drain(bp >> 5);
  }
  __device__ __forceinline__ void drain(uint32_t nwords)
  {
    if (!sink) {
      for (uint32_t j = 0; j < nwords; j++) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (j << 7)));
        gdst[j] = v;
      }
      gdst += nwords;
    }
    drained += nwords * 32;
    bp &= 31;  // the partial word lives in acc; the next append rewrites column word 0 from it
  }
  // single pass: the partial word joins the column, nothing leaves
  __device__ __forceinline__ void finish_window()
  {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(word_addr()), "r"(acc) : "memory");
  }
  // variable rate: everything (including the partial word) to the slot
  __device__ __forceinline__ void finish_slot()
  {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(word_addr()), "r"(acc) : "memory");
    drain((bp + 31) >> 5);
  }
  __device__ __forceinline__ uint32_t word_addr() const { return base + ((bp & ~31u) << 2); }
  __device__ __forceinline__ void append32(uint32_t v, uint32_t len)  // v < 2^len, len <= 32
  {
    const uint32_t sh = bp & 31, w = word_addr();
    // (shift + OR + funnel shift; forming a0:a1 with one IMAD.WIDE on the idle FMA pipe measured slower)
    const uint32_t a0 = acc | (v << sh);
    const uint32_t a1 = __funnelshift_l(v, 0, sh);
    asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+128], %2;" ::"r"(w), "r"(a0), "r"(a1) : "memory");
    acc = sh + len < 32 ? a0 : a1;
    bp += len;
  }
  __device__ __forceinline__ void append64(uint32_t lo, uint32_t hi, uint32_t len)  // (hi:lo) < 2^len, len <= 64
  {
    const uint32_t sh = bp & 31, w = word_addr();
    const uint32_t a0 = acc | (lo << sh);
    const uint32_t a1 = __funnelshift_l(lo, hi, sh);
    const uint32_t a2 = __funnelshift_l(hi, 0, sh);
    asm volatile("st.shared.u32 [%0], %1;\n\tst.shared.u32 [%0+128], %2;\n\tst.shared.u32 [%0+256], %3;" ::"r"(w), "r"(a0), "r"(a1),
                 "r"(a2)
                 : "memory");
    const uint32_t t = sh + len;  // <= 95
    acc = t < 32 ? a0 : (t < 64 ? a1 : a2);
    bp += len;
  }
  __device__ __forceinline__ void put(uint64_t v, uint32_t len) { append64((uint32_t)v, (uint32_t)(v >> 32), len); }
  __device__ __forceinline__ void pad(uint32_t zeros)  // fixed rate: finish() zero-fills instead
  {
    if (gdst) {
      while (zeros) {
        const uint32_t step = zeros < 32 ? zeros : 32;
        append32(0, step);
        zeros -= step;
        drain_if_low();
      }
    }
  }
  __device__ __forceinline__ uint32_t tell() const { return drained + bp; }
  // close the block at exactly total_words 32-bit words: zero-fill; overshoot is simply not copied
  __device__ __forceinline__ void finish(uint32_t total_words)
  {
    for (uint32_t w = (bp >> 5) + 1; w < total_words; w++)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + (w << 7)), "r"(0u) : "memory");
  }
};

// bit reader with a 64-bit look-ahead window
struct BitReader {
  static constexpr bool kStaged = false;
  const uint64_t* w;  // next word to fetch
  uint64_t buf;       // unread bits, LSB first
  uint32_t avail;     // number of valid bits in buf (0..64)
  uint32_t lut4 = 0;  // shared-space address of a copy of kDecLut4 (0: none), blocks of four values (decode_planes)

  __device__ __forceinline__ void init(const void* words, uint64_t bitpos)
  {
    w = reinterpret_cast<const uint64_t*>(words) + (bitpos >> 6);
    uint32_t sh = (uint32_t)(bitpos & 63);
    buf = __ldg(w++) >> sh;
    avail = 64 - sh;
  }
  // next len bits without consuming, 0 <= len <= 64
  __device__ __forceinline__ uint64_t peek(uint32_t len)
  {
    uint64_t v = buf;
    if (len > avail) {
      uint64_t nx = __ldg(w);
      v |= avail < 64 ? nx << avail : 0;
    }
    return v & lowmask64(len);
  }
  __device__ __forceinline__ void skip(uint32_t len)
  {
    if (len < avail) {
      buf >>= len;
      avail -= len;
    }
    else if (len == avail) {
      buf = 0;  // fetch lazily: never touch a word the stream may not own
      avail = 0;
    }
    else {
      uint32_t need = len - avail;  // 1..64
      uint64_t nx = __ldg(w++);
      buf = need < 64 ? nx >> need : 0;
      avail = 64 - need;
    }
  }
  __device__ __forceinline__ uint64_t get(uint32_t len)
  {
    uint64_t v = peek(len);
    skip(len);
    return v;
  }
};

// A block of at most 64 bits held in a register pair (1-D blocks at up to 16 bits per value): no refills.
struct SmallReader {
  static constexpr bool kStaged = false;
  uint64_t buf;       // unread bits, LSB first; zero beyond the block
  uint32_t lut4 = 0;  // as in BitReader

  __device__ __forceinline__ void init(const void* words, uint64_t bitpos, uint32_t nbits)  // nbits <= 64
  {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(words) + (bitpos >> 6);
    const uint32_t sh = (uint32_t)(bitpos & 63);
    buf = __ldg(w) >> sh;
    if (sh + nbits > 64)  // (never touch a word the stream may not own)
      buf |= __ldg(w + 1) << (64 - sh);
    buf &= lowmask64(nbits);
  }
  __device__ __forceinline__ uint64_t peek(uint32_t len) const { return buf & lowmask64(len); }
  __device__ __forceinline__ void skip(uint32_t len) { buf = shr64c(buf, len); }
  __device__ __forceinline__ uint64_t get(uint32_t len)
  {
    const uint64_t v = peek(len);
    skip(len);
    return v;
  }
};

// Column reader (plane-lockstep fixed-rate path): the block's words sit in a lane-private
// [word][lane] column followed by zero words; reads are by bit position straight from shared memory
// (no register window to maintain), so all loads of a plane are issued together.
struct ColReader {
  static constexpr bool kStaged = true;
  static constexpr bool kLockstep = true;
  uint32_t base;  // shared-space byte address of word 0 of this lane's column (words are 128 bytes apart)
  uint32_t bp;    // read position in the column, bits
  // variable rate: the column holds a `cap`-word window of the block's words, which start at gsrc
  // (32-bit words) and number `left` from there; restage() slides the window
  const uint32_t* gsrc;
  uint32_t left, cap;
  // shared-space address of the 32-entry table of test-bit positions used by the run decoder
  // (decode_planes_events): entry n has bits n, 2n+1, 3n+2, ... below 32 set
  uint32_t pm;
  // shared-space address of a copy of kDecLut8h (0: none): the small-universe plane steps, decode_pair_small8
  uint32_t lut8;

  __device__ __forceinline__ void init(const uint32_t* column)
  {
    base = (uint32_t)__cvta_generic_to_shared(column);
    bp = 0;
    gsrc = nullptr;
    left = cap = 0;
    pm = 0;
    lut8 = 0;
  }
  __device__ __forceinline__ void set_run_table(const uint32_t* table) { pm = (uint32_t)__cvta_generic_to_shared(table); }
  // fill the table (one thread per entry; the caller synchronises the CTA afterwards)
  __device__ static __forceinline__ void fill_run_table(uint32_t* table, uint32_t n)
  {
    uint32_t m = 0;
    for (uint32_t pos = n; pos < 32; pos += n + 1)
      m |= 1u << pos;
    table[n] = m;
  }
  __device__ __forceinline__ uint32_t run_mask(uint32_t n) const  // n < 32
  {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(pm + (n << 2)));
    return v;
  }
  __device__ __forceinline__ void stage()
  {
    const uint32_t n = left < cap ? left : cap;
    for (uint32_t j = 0; j < cap; j++)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + (j << 7)), "r"(j < n ? __ldg(gsrc + j) : 0u) : "memory");
  }
  __device__ __forceinline__ void init_var(const uint32_t* column, uint32_t capacity_words, const uint32_t* src, uint32_t words,
                                           uint32_t phase)
  {
    base = (uint32_t)__cvta_generic_to_shared(column);
    cap = capacity_words;
    gsrc = src;
    left = words;
    pm = 0;
    lut8 = 0;
    stage();
    bp = phase;
  }
  // checked every two planes: each reads < 200 bits past its start and every read looks up to three words ahead
  __device__ __forceinline__ bool needs_restage() const { return gsrc && bp > (cap - 16) * 32; }
  __device__ __forceinline__ void restage()
  {
    const uint32_t adv = bp >> 5;
    gsrc += adv;
    left = left > adv ? left - adv : 0;
    stage();
    bp &= 31;
  }
  __device__ __forceinline__ uint32_t peek32(uint32_t pos) const  // the 32 bits starting at bit `pos`
  {
    uint32_t w0, w1;
    asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+128];" : "=r"(w0), "=r"(w1) : "r"(base + ((pos & ~31u) << 2)));
    return __funnelshift_r(w0, w1, pos);  // shift amount is taken modulo 32
  }
  __device__ __forceinline__ void peek64(uint32_t pos, uint32_t& lo, uint32_t& hi) const
  {
    uint32_t w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+128];\n\tld.shared.u32 %2, [%3+256];"
                 : "=r"(w0), "=r"(w1), "=r"(w2)
                 : "r"(base + ((pos & ~31u) << 2)));
    lo = __funnelshift_r(w0, w1, pos);
    hi = __funnelshift_r(w1, w2, pos);
  }
  __device__ __forceinline__ uint64_t get(uint32_t len)  // len <= 64
  {
    uint32_t lo, hi;
    peek64(bp, lo, hi);
    bp += len;
    const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32);
    return v ^ shl64c(shr64c(v, len), len);
  }
};

// ------------------------------------------------------------------------------------------------
// lifting transforms on register-resident values (wrapping arithmetic via unsigned types)
// ------------------------------------------------------------------------------------------------
template <class Int>
__device__ __forceinline__ void fwd_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  x = (Int)((U)x + (U)w); x >>= 1; w = (Int)((U)w - (U)x);
  z = (Int)((U)z + (U)y); z >>= 1; y = (Int)((U)y - (U)z);
  x = (Int)((U)x + (U)z); x >>= 1; z = (Int)((U)z - (U)x);
  w = (Int)((U)w + (U)y); w >>= 1; y = (Int)((U)y - (U)w);
  w = (Int)((U)w + (U)(y >> 1)); y = (Int)((U)y - (U)(w >> 1));
}

template <class Int>
__device__ __forceinline__ void inv_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  y = (Int)((U)y + (U)(w >> 1)); w = (Int)((U)w - (U)(y >> 1));
  y = (Int)((U)y + (U)w); w = (Int)((U)w * 2u - (U)y);
  z = (Int)((U)z + (U)x); x = (Int)((U)x * 2u - (U)z);
  y = (Int)((U)y + (U)z); z = (Int)((U)z * 2u - (U)y);
  w = (Int)((U)w + (U)x); x = (Int)((U)x * 2u - (U)w);
}

template <class Int>
__device__ __forceinline__ void rev_fwd_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  U a = (U)x, b = (U)y, c = (U)z, d = (U)w;
  d -= c; c -= b; b -= a;
  d -= c; c -= b;
  d -= c;
  y = (Int)b; z = (Int)c; w = (Int)d;
}

template <class Int>
__device__ __forceinline__ void rev_inv_lift(Int& x, Int& y, Int& z, Int& w)
{
  using U = typename std::make_unsigned<Int>::type;
  U a = (U)x, b = (U)y, c = (U)z, d = (U)w;
  d += c;
  c += b; d += c;
  b += a; c += b; d += c;
  y = (Int)b; z = (Int)c; w = (Int)d;
}

// KIND 0: fwd_lift, 1: inv_lift, 2: rev_fwd_lift, 3: rev_inv_lift
template <int KIND, class Int>
__device__ __forceinline__ void lift4(Int& x, Int& y, Int& z, Int& w)
{
  if (KIND == 0) fwd_lift(x, y, z, w);
  else if (KIND == 1) inv_lift(x, y, z, w);
  else if (KIND == 2) rev_fwd_lift(x, y, z, w);
  else rev_inv_lift(x, y, z, w);
}

template <int KIND, int DIMS, int AXIS, class Int>
__device__ __forceinline__ void lift_axis(Int (&p)[1 << (2 * DIMS)])
{
  constexpr int N = 1 << (2 * DIMS), st = 1 << (2 * AXIS);
#pragma unroll
  for (int i = 0; i < N; i++)
    if (((i >> (2 * AXIS)) & 3) == 0)
      lift4<KIND>(p[i], p[i + st], p[i + 2 * st], p[i + 3 * st]);
}

// forward: x, y, z (encode{1,2,3}.c fwd_xform); inverse: z, y, x (decode{1,2,3}.c inv_xform)
template <int KIND, int DIMS, class Int>
__device__ __forceinline__ void xform_fwd(Int (&p)[1 << (2 * DIMS)])
{
  lift_axis<KIND, DIMS, 0>(p);
  if (DIMS > 1) lift_axis<KIND, DIMS, (DIMS > 1 ? 1 : 0)>(p);
  if (DIMS > 2) lift_axis<KIND, DIMS, (DIMS > 2 ? 2 : 0)>(p);
}
template <int KIND, int DIMS, class Int>
__device__ __forceinline__ void xform_inv(Int (&p)[1 << (2 * DIMS)])
{
  if (DIMS > 2) lift_axis<KIND, DIMS, (DIMS > 2 ? 2 : 0)>(p);
  if (DIMS > 1) lift_axis<KIND, DIMS, (DIMS > 1 ? 1 : 0)>(p);
  lift_axis<KIND, DIMS, 0>(p);
}

// The inverse transform of 64-bit coefficients on the FP64 pipe (decode only).  A block whose lowest
// decoded plane is L has coefficients that are multiples of 2^L; one inv_lift halves twice
// ("y += w >> 1; w -= y >> 1", decode.c:13-45 of the reference's template), so after d passes every
// intermediate is still a multiple of 2^(L-2d): no shift ever drops a bit and the arithmetic is
// plain linear algebra over the integers.  While nothing leaves the int64 range the values are also
// doubles with at most 63 - (L - 2d) <= 53 significant bits when L >= 10 + 2d: each step is then ONE
// exact DADD / DFMA instead of two to four instructions on the integer pipe, and the result equals
// the integer transform bit for bit (wrapping can only matter where a shift or the final conversion
// looks at a wrapped value; lift_gain() bounds every such value, see decode_block).
__device__ __forceinline__ void inv_lift_f64(double& x, double& y, double& z, double& w)
{
  y = __fma_rn(w, 0.5, y); w = __fma_rn(y, -0.5, w);
  y = __dadd_rn(y, w); w = __fma_rn(w, 2.0, -y);
  z = __dadd_rn(z, x); x = __fma_rn(x, 2.0, -z);
  y = __dadd_rn(y, z); z = __fma_rn(z, 2.0, -y);
  w = __dadd_rn(w, x); x = __fma_rn(x, 2.0, -w);
}
template <int DIMS, int AXIS>
__device__ __forceinline__ void inv_lift_axis_f64(double (&p)[1 << (2 * DIMS)])
{
  constexpr int N = 1 << (2 * DIMS), st = 1 << (2 * AXIS);
#pragma unroll
  for (int i = 0; i < N; i++)
    if (((i >> (2 * AXIS)) & 3) == 0)
      inv_lift_f64(p[i], p[i + st], p[i + 2 * st], p[i + 3 * st]);
}
template <int DIMS>
__device__ __forceinline__ void xform_inv_f64(double (&p)[1 << (2 * DIMS)])
{
  if (DIMS > 2) inv_lift_axis_f64<DIMS, (DIMS > 2 ? 2 : 0)>(p);
  if (DIMS > 1) inv_lift_axis_f64<DIMS, (DIMS > 1 ? 1 : 0)>(p);
  inv_lift_axis_f64<DIMS, 0>(p);
}
// Largest magnitude with which coefficient i (natural order) enters any value the inverse transform
// shifts or returns: per axis the largest entry of its column over the rows y + w/2 (the shifted
// intermediate) and the four outputs (x, y, z, w columns: 1, 3/2, 1, 5/4), multiplied over the axes.
__host__ __device__ constexpr double lift_gain(int i, int dims)
{
  double g = 1;
  for (int d = 0; d < dims; d++) {
    const int j = (i >> (2 * d)) & 3;
    g *= j == 1 ? 1.5 : (j == 3 ? 1.25 : 1.0);
  }
  return g;
}

// ------------------------------------------------------------------------------------------------
// 32x32 bit-matrix transpose in registers: on return bit i of a[j] is the former bit j of a[i]
// ------------------------------------------------------------------------------------------------
// one LOP3 with an explicit truth table (the compiler splits (x & m) | (y & ~m) into two operations
// with two immediates when left to itself).  Table bit i = f(a, b, c) for (a, b, c) = bits 2, 1, 0 of i.
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return r;
}
// (x & m) | (y & ~m), optionally with y or the result inverted
template <int NEG>
__device__ __forceinline__ uint32_t bitselect(uint32_t x, uint32_t y, uint32_t m)
{
  return NEG == 0 ? lop3<0xE4>(x, y, m) : NEG == 2 ? lop3<0xB1>(x, y, m) : lop3<0x1B>(x, y, m);
}
// One butterfly stage: rows k and k+J exchange the column groups selected by m.  Written as two
// bit-selects (one LOP3 each) on the shifted partner instead of the classic xor-swap
// (shift, xor-and, xor, shift, xor): 3 ALU-pipe instructions + 1 left shift that ptxas may place on
// the FMA pipe (IMAD.SHL) per pair instead of 5.
// NEG folds the negabinary mask 0xaaaa... (encode.c:75-88 int2uint, decode.c:63-76 uint2int) into the
// J == 1 stage: XOR with the mask is "invert every odd bit plane", i.e. every odd row on the plane
// side of the transpose, and an inverted input or output costs nothing in a LOP3.
//   NEG 0: plain.   NEG 1: odd OUTPUT rows inverted (coefficients -> planes, J == 1 is the last stage).
//   NEG 2: odd INPUT rows inverted (planes -> coefficients, J == 1 is the first stage).
// (The butterfly stages swap independent index bits, so they commute.)
template <int J, int NEG = 0>
__device__ __forceinline__ void transpose32_stage(uint32_t (&a)[32])
{
  constexpr uint32_t m = J == 16 ? 0x0000ffffu : J == 8 ? 0x00ff00ffu : J == 4 ? 0x0f0f0f0fu : J == 2 ? 0x33333333u : 0x55555555u;
  static_assert(NEG == 0 || J == 1, "the negabinary fold lives in the J == 1 stage");
#pragma unroll
  for (int k = 0; k < 32; k++)
    if (!(k & J)) {
      // swap (row k, columns c+J) with (row k+J, columns c) for the columns c selected by m
      if (J == 16) {
        // halfword granularity: two byte permutes instead of shifts and selects
        const uint32_t lo = __byte_perm(a[k], a[k + J], 0x5410), hi = __byte_perm(a[k], a[k + J], 0x7632);
        a[k] = lo;
        a[k + J] = hi;
      }
      else if (J == 8) {
        const uint32_t lo = __byte_perm(a[k], a[k + J], 0x6240), hi = __byte_perm(a[k], a[k + J], 0x7351);
        a[k] = lo;
        a[k + J] = hi;
      }
      else {
        // NEG 2: row k+J comes in inverted (both selects take ~r1); NEG 1: row k+J goes out inverted
        const uint32_t r0 = a[k], r1 = a[k + J];
        a[k] = bitselect<NEG == 2 ? 2 : 0>(r0, r1 << J, m);
        a[k + J] = bitselect<NEG>(r0 >> J, r1, m);
      }
    }
}

template <int NEG = 0>
__device__ __forceinline__ void transpose32(uint32_t (&a)[32])
{
  if (NEG == 2) transpose32_stage<1, NEG>(a);
  transpose32_stage<16>(a);
  transpose32_stage<8>(a);
  transpose32_stage<4>(a);
  transpose32_stage<2>(a);
  if (NEG != 2) transpose32_stage<1, NEG>(a);
}

// Small blocks (M = 16 or 4 coefficients): the 32-bit words of the M coefficients form 32/M square
// M x M bit matrices side by side; transposing each in place (log2 M butterfly stages on M words,
// the masks are periodic so all matrices go at once) leaves word i = [plane i | plane M+i | ...],
// M bits each (plane parity = row parity, so the negabinary fold works as above).  A quarter / a
// sixteenth of the work of padding to a 32x32 matrix.
template <int J, int M, int NEG = 0>
__device__ __forceinline__ void transpose_small_stage(uint32_t (&a)[M])
{
  constexpr uint32_t m = J == 8 ? 0x00ff00ffu : J == 4 ? 0x0f0f0f0fu : J == 2 ? 0x33333333u : 0x55555555u;
  static_assert(NEG == 0 || J == 1, "the negabinary fold lives in the J == 1 stage");
#pragma unroll
  for (int k = 0; k < M; k++)
    if (!(k & J)) {
      if (J == 8) {
        const uint32_t lo = __byte_perm(a[k], a[k + J], 0x6240), hi = __byte_perm(a[k], a[k + J], 0x7351);
        a[k] = lo;
        a[k + J] = hi;
      }
      else {
        // NEG 2: row k+J comes in inverted (both selects take ~r1); NEG 1: row k+J goes out inverted
        const uint32_t r0 = a[k], r1 = a[k + J];
        a[k] = bitselect<NEG == 2 ? 2 : 0>(r0, r1 << J, m);
        a[k + J] = bitselect<NEG>(r0 >> J, r1, m);
      }
    }
}
template <int M, int NEG = 0>
__device__ __forceinline__ void transpose_small(uint32_t (&a)[M])
{
  if (NEG == 2) transpose_small_stage<1, M, NEG>(a);
  if constexpr (M >= 16) transpose_small_stage<8, M>(a);
  if constexpr (M >= 8) transpose_small_stage<4, M>(a);
  transpose_small_stage<2, M>(a);
  if (NEG != 2) transpose_small_stage<1, M, NEG>(a);
}

// Plane storage in shared memory: plane k of the calling lane is sp[k * 32] where sp already
// points at the lane's column (conflict free for 4- and 8-byte words).
template <int N> struct PlaneWord { using type = uint32_t; };
template <> struct PlaneWord<64> { using type = uint64_t; };

// In all plane conversions NEG says where the negabinary mask is (see transpose32_stage): with
// NEG = 1 the input words are "pre-negabinary" coefficients q + 0xaaaa... and the planes come out as
// those of (q + mask) ^ mask; with NEG = 2 the planes go in as stored and the words come out as
// u ^ 0xaaaa..., to which the caller applies the subtraction (planes that were never decoded read
// as zero, so their bits come out as the mask's: u ^ mask for u = 0).  NEG = 0 (reversible mode,
// whose precision scan needs the true negabinary words) leaves both sides alone.
template <int NEG> struct NegaWord {
  static constexpr uint32_t w32 = NEG ? 0xaaaaaaaau : 0u;        // what a word of the NEG-side representation holds for u = 0
  static constexpr uint64_t w64 = NEG ? 0xaaaaaaaaaaaaaaaaull : 0ull;
};

// coefficients (sequency order) -> bit planes
template <int NEG, class UInt, int N, int STRIDE = 32>
__device__ __forceinline__ void to_planes(const UInt (&u)[N], typename PlaneWord<N>::type* sp)
{
  constexpr int P = 8 * (int)sizeof(UInt);
  constexpr int G = (N + 31) / 32;  // groups of 32 coefficients (1 or 2)
#pragma unroll
  for (int h = 0; h < P / 32; h++) {
    uint32_t a[G][32];
#pragma unroll
    for (int g = 0; g < G; g++) {
#pragma unroll
      for (int i = 0; i < 32; i++)
        a[g][i] = (32 * g + i < N) ? (uint32_t)(u[(32 * g + i) % N] >> (32 * h)) : NegaWord<NEG>::w32;
      transpose32<NEG>(a[g]);
    }
#pragma unroll
    for (int k = 0; k < 32; k++) {
      if (G == 2)
        sp[(32 * h + k) * STRIDE] = (typename PlaneWord<N>::type)((uint64_t)a[0][k] | ((uint64_t)a[G - 1][k] << 32));
      else
        sp[(32 * h + k) * STRIDE] = (typename PlaneWord<N>::type)a[0][k];
    }
  }
}

// bit planes -> coefficients; planes below kstop were never written and read as zero
template <int NEG, class UInt, int N, int STRIDE = 32>
__device__ __forceinline__ void from_planes(UInt (&u)[N], const typename PlaneWord<N>::type* sp, int kstop)
{
  constexpr int P = 8 * (int)sizeof(UInt);
  constexpr int G = (N + 31) / 32;
#pragma unroll
  for (int i = 0; i < N; i++)
    u[i] = 0;
#pragma unroll
  for (int h = 0; h < P / 32; h++) {
    uint32_t a[G][32];
#pragma unroll
    for (int k = 0; k < 32; k++) {
      typename PlaneWord<N>::type x = (32 * h + k >= kstop) ? sp[(32 * h + k) * STRIDE] : 0;
      a[0][k] = (uint32_t)x;
      if (G == 2)
        a[G - 1][k] = (uint32_t)((uint64_t)x >> 32);
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
      transpose32<NEG>(a[g]);
#pragma unroll
      for (int i = 0; i < 32; i++)
        if (32 * g + i < N)
          u[(32 * g + i) % N] |= (UInt)((UInt)a[g][i] << (32 * h));
    }
  }
}

// 32-bit half H of word u replaced by v
template <int H> __device__ __forceinline__ void set_half(uint32_t& u, uint32_t v) { u = v; }
template <int H> __device__ __forceinline__ void set_half(uint64_t& u, uint32_t v)
{
  u = H ? ((u & 0xffffffffull) | ((uint64_t)v << 32)) : ((u & 0xffffffff00000000ull) | v);
}

// Two-phase variants for the staged fast path: only 32 planes (bits 32H .. 32H+31 of every
// coefficient) are resident in shared memory at a time, as sp[(k - 32H) * 32].  Halves the plane
// storage (more resident warps) and skips the low half entirely when no block of the warp needs it.
// Blocks of 64 values: returns `kup`, the number of planes at the bottom of this set (0..32) in which
// some block of the warp has a bit beyond coefficient 31 (one warp-wide OR of the words of
// coefficients 32..63; on smooth data the high-sequency coefficients are small and only the lowest
// coded planes reach them).  Planes kup.. of the set are 32-bit work for the coder (narrow plane
// steps), and with kup == 0 the words of coefficients 32..63 are not transposed at all.
template <int H, int NEG, class UInt, int N>
__device__ __forceinline__ int to_planes_half(const UInt (&u)[N], typename PlaneWord<N>::type* sp, int* ksmall8 = nullptr)
{
  if constexpr (N == 16 || N == 4) {
    uint32_t a[N];
#pragma unroll
    for (int i = 0; i < N; i++)
      a[i] = (uint32_t)(u[i] >> (32 * H));
    if constexpr (N == 16) {
      if (ksmall8) {
        uint32_t any_upper = 0;
#pragma unroll
        for (int i = 8; i < 16; i++)
          any_upper |= a[i] ^ NegaWord<NEG>::w32;
        *ksmall8 = 32 - __clz((int)__reduce_or_sync(0xffffffffu, any_upper));
      }
    }
    transpose_small<N, NEG>(a);
#pragma unroll
    for (int q = 0; q < 32 / N; q++)
#pragma unroll
      for (int i = 0; i < N; i++)
        sp[(N * q + i) * 32] = (a[i] >> (N * q)) & ((1u << N) - 1);
    return 0;
  }
  else {
    static_assert(N == 64, "blocks of 4, 16 or 64 values");
    uint32_t a0[32], a1[32];
    uint32_t any_upper = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
      a0[i] = (uint32_t)(u[i] >> (32 * H));
      a1[i] = (uint32_t)(u[32 + i] >> (32 * H));
      any_upper |= a1[i] ^ NegaWord<NEG>::w32;
    }
    const int kup = 32 - __clz((int)__reduce_or_sync(0xffffffffu, any_upper));
    if (ksmall8) {
      // planes >= *ksmall8 of the set have one-bits in coefficients 0..7 only, in every block of the warp
      uint32_t any_mid = any_upper;
#pragma unroll
      for (int i = 8; i < 32; i++)
        any_mid |= a0[i] ^ NegaWord<NEG>::w32;
      *ksmall8 = 32 - __clz((int)__reduce_or_sync(0xffffffffu, any_mid));
    }
    transpose32<NEG>(a0);
    if (kup == 0) {
#pragma unroll
      for (int k = 0; k < 32; k++)
        sp[k * 32] = (uint64_t)a0[k];
    }
    else {
      transpose32<NEG>(a1);
#pragma unroll
      for (int k = 0; k < 32; k++)
        sp[k * 32] = (uint64_t)a0[k] | ((uint64_t)a1[k] << 32);
    }
    return kup;
  }
}

// Planes 32H .. 32H+31 (those with absolute index >= kstop; the rest read as zero) replace half H of
// every word of u.  `upper` = false: no block of the warp has a significant coefficient beyond the
// 32nd, so the words of coefficients 32..63 keep the value they were initialised with (u = 0).
template <int H, int NEG, class UInt, int N>
__device__ __forceinline__ void from_planes_half(UInt (&u)[N], const typename PlaneWord<N>::type* sp, int kstop, bool upper = true)
{
  if constexpr (N == 16 || N == 4) {
    uint32_t a[N];
#pragma unroll
    for (int i = 0; i < N; i++)
      a[i] = 0;
#pragma unroll
    for (int q = 0; q < 32 / N; q++)
#pragma unroll
      for (int i = 0; i < N; i++) {
        const uint32_t x = (32 * H + N * q + i >= kstop) ? (uint32_t)sp[(N * q + i) * 32] : 0u;
        a[i] |= x << (N * q);
      }
    transpose_small<N, NEG>(a);
#pragma unroll
    for (int i = 0; i < N; i++)
      set_half<H>(u[i], a[i]);
  }
  else {
    static_assert(N == 64, "blocks of 4, 16 or 64 values");
    uint32_t a0[32], a1[32];
#pragma unroll
    for (int k = 0; k < 32; k++) {
      const uint64_t x = (32 * H + k >= kstop) ? sp[k * 32] : 0;
      a0[k] = (uint32_t)x;
      a1[k] = (uint32_t)(x >> 32);
    }
    transpose32<NEG>(a0);
#pragma unroll
    for (int i = 0; i < 32; i++)
      set_half<H>(u[i], a0[i]);
    if (upper) {
      transpose32<NEG>(a1);
#pragma unroll
      for (int i = 0; i < 32; i++)
        set_half<H>(u[32 + i], a1[i]);
    }
  }
}

// 16-plane windows for 64 coefficients of 64 bits (3-D blocks of double / int64): planes
// 16W .. 16W+15 are resident as sp[(k - 16W) * 32].  Rows r = 16 rb + l of the bit matrix hold the
// 16-bit slices of coefficients 32 rb + l (low half) and 32 rb + 16 + l (high half); transposing
// every 16x16 block in place (the last four butterfly stages - the halfword stage is what the
// packing already did) leaves row 16 rb + i = plane i of coefficients 32 rb .. 32 rb + 31 (plane
// parity = row parity).  Same cost per plane as the 32-plane halves, but blocks that stop a few planes
// into the low half - the usual case at rates around 8 - pay for 16 more planes instead of 32.
// The window is 16-bit slice w (0 or 1, a run-time value: it only changes a byte-permute selector)
// of 32-bit half H (compile time: it selects registers), so the coder loop that follows is
// instantiated once per half, not once per window.  Returns kup (0..16) as to_planes_half does; rows
// 16..31 (coefficients 32..63) are skipped when it is 0.
template <int J, int NEG>
__device__ __forceinline__ void transpose16_stage(uint32_t (&a)[16])
{
  constexpr uint32_t m = J == 8 ? 0x00ff00ffu : J == 4 ? 0x0f0f0f0fu : J == 2 ? 0x33333333u : 0x55555555u;
#pragma unroll
  for (int k = 0; k < 16; k++)
    if (!(k & J)) {
      if (J == 8) {
        const uint32_t lo = __byte_perm(a[k], a[k + J], 0x6240), hi = __byte_perm(a[k], a[k + J], 0x7351);
        a[k] = lo;
        a[k + J] = hi;
      }
      else {
        // NEG 2: row k+J comes in inverted (both selects take ~r1); NEG 1: row k+J goes out inverted
        const uint32_t r0 = a[k], r1 = a[k + J];
        a[k] = bitselect<NEG == 2 ? 2 : 0>(r0, r1 << J, m);
        a[k + J] = bitselect<NEG>(r0 >> J, r1, m);
      }
    }
}
// two 16x16 bit matrices side by side (low / high halfword of 16 words), each transposed in place
template <int NEG>
__device__ __forceinline__ void transpose16x2(uint32_t (&a)[16])
{
  if (NEG == 2) transpose16_stage<1, NEG>(a);
  transpose16_stage<8, 0>(a);
  transpose16_stage<4, 0>(a);
  transpose16_stage<2, 0>(a);
  if (NEG != 2) transpose16_stage<1, NEG>(a);
}

template <int H, int NEG>
__device__ __forceinline__ int to_planes_window(const uint64_t (&u)[64], uint64_t* sp, uint32_t w)
{
  const uint32_t sel = w ? 0x7632u : 0x5410u;
  uint32_t a0[16], a1[16];
  uint32_t any_upper = 0;
#pragma unroll
  for (int l = 0; l < 16; l++) {
    a0[l] = __byte_perm((uint32_t)(u[l] >> (32 * H)), (uint32_t)(u[16 + l] >> (32 * H)), sel);
    a1[l] = __byte_perm((uint32_t)(u[32 + l] >> (32 * H)), (uint32_t)(u[48 + l] >> (32 * H)), sel);
    any_upper |= a1[l] ^ NegaWord<NEG>::w32;
  }
  const uint32_t any_all = __reduce_or_sync(0xffffffffu, any_upper);  // two 16-bit slices per word
  const int kup = 32 - __clz((int)((any_all | (any_all >> 16)) & 0xffffu));
  transpose16x2<NEG>(a0);
  if (kup == 0) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      sp[i * 32] = (uint64_t)a0[i];
  }
  else {
    transpose16x2<NEG>(a1);
#pragma unroll
    for (int i = 0; i < 16; i++)
      sp[i * 32] = (uint64_t)a0[i] | ((uint64_t)a1[i] << 32);
  }
  return kup;
}

// Planes 16W .. 16W+15 (those with absolute index >= kstop; the rest read as zero) replace bits
// 16W .. 16W+15 of every word of u.  Compile-time window: the decoder measured faster with one
// coder-loop instance per window (1024^3 fp64 rate 8: 6.1 ms against 6.4 ms), the encoder with one
// per half.  `upper` as in from_planes_half.
template <int W, int NEG>
__device__ __forceinline__ void from_planes_window(uint64_t (&u)[64], const uint64_t* sp, int kstop, bool upper = true)
{
  uint32_t a0[16], a1[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const uint64_t x = (16 * W + i >= kstop) ? sp[i * 32] : 0;
    a0[i] = (uint32_t)x;
    a1[i] = (uint32_t)(x >> 32);
  }
  // 16-bit slice W of the low word of u replaced by the low (coefficient l) / high (16 + l) halfword of a
  constexpr uint32_t sel_lo = W ? 0x1054u : 0x7610u, sel_hi = W ? 0x3254u : 0x7632u;  // __byte_perm(a, u_lo, sel): bytes 0-3 = a, 4-7 = u_lo
  transpose16x2<NEG>(a0);
#pragma unroll
  for (int l = 0; l < 16; l++) {
    u[l] = (u[l] & 0xffffffff00000000ull) | __byte_perm(a0[l], (uint32_t)u[l], sel_lo);
    u[16 + l] = (u[16 + l] & 0xffffffff00000000ull) | __byte_perm(a0[l], (uint32_t)u[16 + l], sel_hi);
  }
  if (upper) {
    transpose16x2<NEG>(a1);
#pragma unroll
    for (int l = 0; l < 16; l++) {
      u[32 + l] = (u[32 + l] & 0xffffffff00000000ull) | __byte_perm(a1[l], (uint32_t)u[32 + l], sel_lo);
      u[48 + l] = (u[48 + l] & 0xffffffff00000000ull) | __byte_perm(a1[l], (uint32_t)u[48 + l], sel_hi);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// embedded coder on plane words (N <= 64)
// ------------------------------------------------------------------------------------------------

// Emit planes P-1 .. P-maxprec truncated at `budget` bits; returns the bits used.
template <int N, int P, class Writer>
__device__ __forceinline__ uint32_t encode_planes(Writer& bw, uint32_t budget, uint32_t maxprec,
                                                  const typename PlaneWord<N>::type* sp)
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  for (int k = P - 1; bits && k >= kmin; k--) {
    uint64_t x = sp[k * 32];
    // bits of the n coefficients already significant, verbatim
    uint32_t m = n < bits ? n : bits;
    bits -= m;
    bw.put(x & lowmask64(m), m);
    // remaining coefficients: group test, then the distance to the next one-bit in unary
    uint64_t r = n < 64 ? x >> n : 0;
    while (bits && n < N) {
      if (!r) {
        bw.put(0, 1);
        bits--;
        break;
      }
      uint32_t c = ctz64(r);
      bool implied = n + c == N - 1;  // a one-bit in the last slot is not written
      uint32_t len = implied ? c + 1 : c + 2;
      uint64_t val = implied ? 1ull : (1ull | (2ull << c));
      if (len > bits) {
        len = bits;
        val &= lowmask64(len);
      }
      bw.put(val, len);
      bits -= len;
      n += c + 1;
      r >>= c;
      r >>= 1;
    }
  }
  return budget - bits;
}

template <class T> __device__ __forceinline__ uint32_t ctz_any(T x);
template <> __device__ __forceinline__ uint32_t ctz_any<uint32_t>(uint32_t x) { return (uint32_t)__ffs((int)x) - 1; }
template <> __device__ __forceinline__ uint32_t ctz_any<uint64_t>(uint64_t x) { return ctz64(x); }

// Plane-lockstep coder for the column writer.  All 32 lanes walk the planes together (k is warp
// uniform); per plane a lane appends the n verbatim bits in one go, then the whole group-tested
// part T as one word.  T is the region y = x >> n with a flag bit inserted after every one-bit
// ("is there another one above?") behind a leading test bit; the last flag is the plane's closing
// '0' test, and a one-bit on the last coefficient is implied together with its flag
// (encode.c:108-124).  With Y' = y's one-bits each moved up by its rank, the doubled string is
// Y' | Y' << 1 = 3 Y'; Y' is built by peeling the lowest one-bit per step with a warp-uniform shift
// count.  Planes whose T does not fit 32 bits (noisy data) take a per-lane run loop instead.
struct LockState {
  uint32_t pos;   // coefficients settled so far = verbatim count of the next plane
  int k;          // last plane coded (starts at P), warp uniform
  bool done;      // budget exhausted or below the block's precision
};

template <int N>
__device__ __forceinline__ void encode_plane_lockstep(ColWriter& bw, uint32_t limit, int kmin, int k, uint32_t& pos, bool& done,
                                                      const typename PlaneWord<N>::type x)
{
  using R = typename PlaneWord<N>::type;
  constexpr uint32_t FULL = 0xffffffffu;
  bw.drain_if_low();
  done = done || k < kmin || bw.tell() >= limit;
  if constexpr (N == 4) {
    // blocks of four values: the plane's whole string by table (kEncLut4, every plane word and every n: no
    // special cases, no votes); a finished lane walks the idle row (empty string)
    if (bw.lut) {
      uint32_t e;
      asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(bw.lut + ((done ? 5u : pos) << 6) + ((uint32_t)x << 2)));
      bw.append32(e & 0xffu, (e >> 8) & 15u);
      pos = e >> 12;
      return;
    }
  }
  const uint32_t n = done ? 0u : pos;  // a finished lane appends nothing: zero-length verbatim part, empty T
  R r, verb;
  if constexpr (N > 32) {
    r = shr64c(x, n);
    verb = x ^ shl64c(r, n);
  }
  else {
    r = shr32c(x, n);
    verb = x ^ shl32c(r, n);
  }
  const R y = r;  // finished lanes: T is forced empty below (has / slow are gated), no selects here
  const bool test = !done && n < N;
  const uint32_t y32 = (uint32_t)y;
  const uint32_t c = (uint32_t)__popc(y32);
  const int msb = 31 - __clz((int)y32);  // -1 when y32 == 0
  // one-bits of the region above its lowest 1, 2, 3, 4 ones
  const uint32_t r1 = y32 & (y32 - 1), r2 = r1 & (r1 - 1), r3 = r2 & (r2 - 1), r4 = r3 & (r3 - 1);
  bool slow = false;
  if constexpr (N > 32)
    slow = !done && ((uint32_t)((uint64_t)y >> 32) != 0 || msb + (int)c + 2 > 32);
  if constexpr (N > 32)
    bw.append64((uint32_t)verb, (uint32_t)((uint64_t)verb >> 32), n);  // (blocks of <= 32 values: appended with T below)
  // every one-bit moved up by its rank: b0 + 2 b1 + 4 b2 + 8 b3 + ... = y + r1 + 2 r2 + 4 r3 + ...
  uint32_t yp = y32 + r1 + 2 * r2 + 4 * r3;
  // one vote on the common path: more than four new coefficients, or a T that does not fit 32 bits
  bool fast = true;
  if (__any_sync(FULL, slow || (!done && r4 != 0))) {
    const uint32_t r5 = r4 & (r4 - 1), r6 = r5 & (r5 - 1), r7 = r6 & (r6 - 1), r8 = r7 & (r7 - 1);
    yp += 8 * r4 + 16 * r5 + 32 * r6 + 64 * r7;
    fast = !__any_sync(FULL, slow || (!done && r8 != 0));
  }
  if (fast) {
    const bool has = !done && y32 != 0;
    const uint32_t top = n + (uint32_t)(msb + 1);    // coefficients settled after this plane
    const uint32_t last = (has && top == N) ? 1u : 0u;
    const uint32_t keep = (uint32_t)msb + c - last;   // bits + flags, minus the closing flag (and the implied pair)
    const uint32_t e = (yp * 3u) & mask32(keep);
    const uint32_t tval = has ? (1u | (e << 1)) : 0u;
    const uint32_t tlen = has ? keep + 2 - last : (test ? 1u : 0u);
    pos = has ? top : pos;
    if constexpr (N > 32)
      bw.append32(tval, tlen);
    else if (!__any_sync(FULL, n + tlen > 32))
      bw.append32((uint32_t)verb | shl32c(tval, n), n + tlen);  // small blocks: the whole plane string in one append
    else {
      bw.append32((uint32_t)verb, n);
      bw.append32(tval, tlen);
    }
  }
  else {
    if constexpr (N <= 32)
      bw.append32((uint32_t)verb, n);
    if (test) {
      R rr = y;
      uint32_t p = n;
      while (rr) {
        const uint32_t z = ctz_any<R>(rr);
        const uint32_t p1 = p + z + 1;
        const uint64_t ex = p1 < N ? 1u : 0u;           // the one-bit is implied on the last coefficient
        const uint64_t v = 1ull | shl64c(ex, z + 1);    // '1', z zeros, '1'
        bw.append64((uint32_t)v, (uint32_t)(v >> 32), z + 1 + (uint32_t)ex);
        rr = (R)shr64c((uint64_t)rr, z + 1);
        p = p1;
      }
      if (p < N)
        bw.append32(0, 1);  // closing (or only) group test
      pos = p;
    }
  }
}

// Narrow plane steps (blocks of 64 values), two planes per call.  Precondition, established by the
// caller for every block of the warp: neither plane word has a bit beyond coefficient 31 and
// pos <= 32.  Then a plane is 32-bit arithmetic: its n verbatim bits and T form one string of at most
// 64 bits, and the formulas need no case split on an empty region: with y == 0, msb = -1 and c = 0
// give tlen = msb + c + 2 = 1 (the lone '0' test), T = 0 and top = n.  A finished lane codes an empty
// plane word with n = 0 and a zero-length string.  Returns false - with nothing appended and no state
// changed - when some lane has more than four new coefficients in a plane or a T longer than 32 bits;
// the caller then runs the general steps on this pair of planes.
struct NarrowPlane {
  uint32_t lo, hi, len, top;  // the plane string (len <= 64) and the coefficients settled after it
  bool fits;
};
__device__ __forceinline__ NarrowPlane narrow_plane_string(bool d, uint32_t pos, uint32_t x)
{
  NarrowPlane r;
  const uint32_t n = d ? 0u : pos;
  const uint32_t xg = d ? 0u : x;
  const uint32_t y = shr32c(xg, n);                 // n == 32: empty region
  const uint32_t verb = xg ^ shl32c(y, n);
  const uint32_t c = (uint32_t)__popc(y);
  const int msb = 31 - __clz((int)y);                // -1 when y == 0
  const uint32_t r1 = y & (y - 1), r2 = r1 & (r1 - 1), r3 = r2 & (r2 - 1), r4 = r3 & (r3 - 1);
  const int keep = msb + (int)c;                     // data bits + flags of T, minus the leading test and the closing flag
  r.fits = r4 == 0 && keep <= 30;
  const uint32_t yp = y + r1 + 2 * r2 + 4 * r3;      // every one-bit moved up by its rank
  const uint32_t e = (yp * 3u) & mask32((uint32_t)keep);  // (keep = -1: y == 0 and e == 0 whatever the mask)
  const uint32_t tval = 2 * e + (y != 0 ? 1u : 0u);
  const uint32_t tlen = d ? 0u : (uint32_t)(keep + 2);
  r.lo = verb | shl32c(tval, n);
  r.hi = __funnelshift_lc(tval, 0u, n);              // n <= 32, tlen <= 32
  r.len = n + tlen;
  r.top = n + (uint32_t)(msb + 1);
  return r;
}

template <int N>
__device__ __forceinline__ bool encode_pair_narrow(ColWriter& bw, uint32_t limit, int kmin, int k, uint32_t& pos, bool& done,
                                                   const uint32_t x1, const uint32_t x2)
{
  bw.drain_if_low();
  const uint32_t at = bw.tell();
  const bool d1 = done || k - 1 < kmin || at >= limit;
  const NarrowPlane p1 = narrow_plane_string(d1, pos, x1);
  const bool d2 = d1 || k - 2 < kmin || at + p1.len >= limit;
  const NarrowPlane p2 = narrow_plane_string(d2, p1.top, x2);
  if (__any_sync(0xffffffffu, !(p1.fits && p2.fits)))
    return false;
  bw.append64(p1.lo, p1.hi, p1.len);
  bw.append64(p2.lo, p2.hi, p2.len);
  done = d2;
  pos = p2.top;  // (a finished lane's pos is never looked at again)
  return true;
}

// Small-universe plane steps (blocks of 16 or 64 values, at most 8 coefficients significant so far): planes
// kbase + 31 down to kbase + ksmall of the resident set have one-bits in coefficients 0..7 only, in every block
// of the warp, so a plane's whole string is one look-up in kEncLut8 by (n, byte) and one append - no special
// cases, and one vote per two planes: the walk stops while every lane still has room for two strings of 17 bits,
// and at the highest precision limit of the warp; the general steps take over at st.k (even).  A finished lane
// walks the table's idle row.
template <int N>
__device__ __forceinline__ void encode_planes_small8(ColWriter& bw, uint32_t limit, int kmin, int kbase, int ksmall, LockState& st,
                                                     const typename PlaneWord<N>::type* sp)
{
  constexpr uint32_t FULL = 0xffffffffu;
  int kstop = kbase + ksmall;
  const int kprec = (int)__reduce_max_sync(FULL, st.done ? 0u : (uint32_t)kmin);
  kstop = kstop > kprec ? kstop : kprec;
  kstop = kstop > kbase ? kstop : kbase;
  kstop = (kstop + 1) & ~1;
  const uint32_t planes = (uint32_t)__cvta_generic_to_shared(sp);
  uint32_t row = bw.lut + ((st.done ? 9u : st.pos) << 10);  // byte address of the table row of n
  int k = st.k;
  for (; k > kstop; k -= 2) {
    if (__any_sync(FULL, !st.done && limit - bw.tell() < 34u))  // (tell() <= limit: the walk never exhausts a budget)
      break;
#pragma unroll
    for (int j = 1; j <= 2; j++) {
      uint32_t b, e;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(planes + (uint32_t)(k - j - kbase) * 32u * (uint32_t)sizeof(typename PlaneWord<N>::type)));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(row + (b << 2)));
      bw.append32(e & 0x1ffffu, (e >> 17) & 31u);
      row = bw.lut + ((e >> 22) << 10);
    }
  }
  st.pos = st.done ? st.pos : (row - bw.lut) >> 10;
  st.k = k;
}

// planes k-1 .. klo of the resident set (first plane kbase), two per vote (k and klo are even).
// knarrow (warp uniform, blocks of 64 values): planes >= knarrow of the set have no bit beyond
// coefficient 31 in any block of the warp (to_planes_half / to_planes_window); together with
// pos <= 32 everywhere that selects the narrow plane steps.
template <int N>
__device__ __forceinline__ void encode_planes_lockstep(ColWriter& bw, uint32_t limit, int kmin, int klo, int kbase,
                                                       LockState& st, const typename PlaneWord<N>::type* sp, int knarrow = 1 << 30)
{
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t pos = st.pos;
  bool done = st.done;
  int k = st.k;
  if constexpr (N > 32 && ZB_NARROW) {
    if (__any_sync(FULL, pos > 32))
      knarrow = 1 << 30;
  }
  while (k > klo && __any_sync(FULL, !done)) {
    const typename PlaneWord<N>::type x1 = sp[(k - 1 - kbase) * 32], x2 = sp[(k - 2 - kbase) * 32];
    if constexpr (N > 32 && ZB_NARROW) {
      if (k - 2 >= knarrow && encode_pair_narrow<N>(bw, limit, kmin, k, pos, done, (uint32_t)x1, (uint32_t)x2)) {
        k -= 2;
        continue;
      }
    }
    encode_plane_lockstep<N>(bw, limit, kmin, k - 1, pos, done, x1);
    encode_plane_lockstep<N>(bw, limit, kmin, k - 2, pos, done, x2);
    k -= 2;
  }
  st.pos = pos;
  st.done = done;
  st.k = k;
}

// Mirror image.  Decoded planes are stored to sp[k*32]; returns bits consumed and, through
// kstop, the lowest plane index that was written.
template <int N, int P, class Reader, bool STORE = true>
__device__ __forceinline__ uint32_t decode_planes(Reader& br, uint32_t budget, uint32_t maxprec,
                                                  typename PlaneWord<N>::type* sp, int& kstop)
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  int k = P - 1;
  for (; bits && k >= kmin; k--) {
    uint32_t m = n < bits ? n : bits;
    bits -= m;
    uint64_t x = br.get(m);
    while (bits && n < N) {
      bits--;
      if (!br.get(1))
        break;
      // scan for the next one-bit: at most L more bits may be read
      uint32_t L = N - 1 - n;
      if (bits < L) L = bits;
      uint64_t t = br.peek(L);
      uint32_t c = t ? ctz64(t) : L;
      uint32_t used = t ? c + 1 : L;
      br.skip(used);
      bits -= used;
      n += c;
      x |= 1ull << n;  // deposited even if the scan ran dry (decode.c:103-111)
      n++;
    }
    if constexpr (STORE)
      sp[k * 32] = (typename PlaneWord<N>::type)x;
  }
  kstop = k + 1;
  return budget - bits;
}

// Blocks of four values with the table (br.lut4): the planes are not stored and transposed afterwards (64 plane
// words and two 32 x 32 transposes for at most 4 x 64 bits) - each decoded plane's four bits go straight into the
// four coefficients.  Returns the bits consumed; u = coefficients ^ NegaWord<NEG> like from_planes<NEG>.
template <int P, int NEG, class UInt, class Reader>
__device__ __forceinline__ uint32_t decode_planes4_direct(Reader& br, uint32_t budget, uint32_t maxprec, UInt (&u)[4])
{
  const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
  uint32_t bits = budget, n = 0;
  // (32 bits of the coefficients at a time: the planes of the upper half first)
  uint32_t c[P / 32][4] = {};
#pragma unroll
  for (int h = P / 32 - 1; h >= 0; h--) {
    uint32_t kbit = 0x80000000u;
    const int klo = 32 * h > kmin ? 32 * h : kmin;
    for (int k = 32 * h + 31; bits && k >= klo; k--, kbit >>= 1) {
      const uint32_t m = n < bits ? n : bits, left = bits - m, a = left < 7 ? left : 7;
      const uint32_t w = (uint32_t)br.peek(m + a);
      const uint32_t ones = (1u << a) - 1;
      uint32_t e;
      asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(br.lut4 + ((5u * ones + (n << a) + ((w >> m) & ones)) << 1)));
      const uint32_t used = m + (e & 15u);
      br.skip(used);
      bits -= used;
      n = e >> 8;
      const uint32_t x = (w & ((1u << m) - 1)) | ((e >> 4) & 15u);
#pragma unroll
      for (int i = 0; i < 4; i++)
        c[h][i] |= (x >> i) & 1u ? kbit : 0u;
    }
  }
  constexpr UInt mask = (UInt)NegaWord<NEG>::w64;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    UInt v = (UInt)c[0][i];
    if constexpr (P == 64)
      v |= (UInt)((uint64_t)c[P / 32 - 1][i] << 32);
    u[i] = v ^ mask;
  }
  return budget - bits;
}

// Plane-lockstep decoder for the column reader, the mirror image of encode_planes_lockstep.  Per
// plane a lane reads its m = min(n, bits) verbatim bits in one go, then parses the whole
// group-tested part T from one 32-bit look at the stream without a loop over its items: with a
// virtual one-bit in front, T is a sequence of runs of ones in which data bits and flag bits
// alternate, starting with a data bit, and T ends after the first run of odd length (its last
// data bit is followed by a '0' flag).  Run ends come from the carries of W + (run starts), split
// by the parity of the start position; the data bits are then moved down by 2 + rank to give the
// region y.  Anything the shortcut cannot prove exact - T longer than the window, the bit budget
// running out inside T, a scan that reaches the last coefficient (implied one-bit, decode.c:103-111)
// - sends the warp through the exact per-item loop for that plane.
struct LockDecodeState {
  uint32_t bits;  // budget left
  uint32_t n;     // coefficients significant so far
  int k;          // last plane decoded (starts at P), warp uniform
  int lowest;     // lowest plane stored by this lane (P when none)
  bool done;
};

// What is left to do for a parsed plane: fetch its m verbatim bits (they start at bp0), merge the
// newly significant bits and store the plane word.  Deferred by one plane so that this independent
// work sits in the same basic block as the next plane's dependent parse chain.
template <int N>
struct PlaneRec {
  typename PlaneWord<N>::type ybits;  // bits decoded from the group-tested part, at their coefficient positions
  uint32_t bp0, m;
  typename PlaneWord<N>::type* dst;
  bool store;
};

template <int N>
__device__ __forceinline__ void finish_plane(const ColReader& br, const PlaneRec<N>& rec)
{
  using R = typename PlaneWord<N>::type;
  R verb;
  if constexpr (N > 32) {
    uint32_t lo, hi;
    br.peek64(rec.bp0, lo, hi);
    const uint64_t v = (uint64_t)lo | ((uint64_t)hi << 32);
    verb = v ^ shl64c(shr64c(v, rec.m), rec.m);
  }
  else {
    const uint32_t v = br.peek32(rec.bp0);
    verb = v ^ shl32c(shr32c(v, rec.m), rec.m);
  }
  if (rec.store)
    *rec.dst = verb | rec.ybits;
}

template <int N>
__device__ __forceinline__ void decode_plane_lockstep(ColReader& br, int kmin, int k, uint32_t& bits, uint32_t& n, int& lowest,
                                                      bool& done, typename PlaneWord<N>::type* plane, PlaneRec<N>& rec)
{
  using R = typename PlaneWord<N>::type;
  constexpr uint32_t FULL = 0xffffffffu;
  const PlaneRec<N> prev = rec;
  done = done || k < kmin || bits == 0;
  const uint32_t m = done ? 0u : (n < bits ? n : bits);  // verbatim bits
  const uint32_t tp = br.bp + m;                         // where T starts
  rec.bp0 = br.bp;
  rec.m = m;
  rec.dst = plane;
  rec.store = !done;
  lowest = done ? lowest : k;
  uint32_t w;
  R verb_small = 0;  // blocks of <= 32 values: verbatim bits and T come from one 64-bit look (m <= 32 - so no deferred fetch)
  if constexpr (N > 32)
    w = br.peek32(tp);
  else {
    uint32_t lo, hi;
    br.peek64(br.bp, lo, hi);
    verb_small = lo ^ shl32c(shr32c(lo, m), m);
    w = m < 32 ? __funnelshift_r(lo, hi, m) : hi;
  }
  const uint32_t left = bits - m;                        // budget at T (m <= bits)
  const bool test = !done && n < N && left != 0;
  // parse T: W' = virtual data bit, then the stream
  const uint32_t wv = (w << 1) | 1u;
  const uint32_t s = wv & ~(wv << 1);                    // run starts
  const uint32_t ae = wv + (s & 0x55555555u), ao = wv + (s & 0xaaaaaaaau);
  const uint32_t term = (ae & ~wv & 0xaaaaaaaau) | (ao & ~wv & 0x55555555u);  // just past each odd-length run
  const uint32_t tl = term & (0u - term);                // lowest end mark; tl - 1 masks the bits of T
  const uint32_t tpos = 31u - (uint32_t)__clz((int)tl);  // bits of T (0xffffffff when no end in the window)
  // data bits of T (virtual one dropped); nothing when there is no group test
  const uint32_t d = ((wv & ~ae & 0x55555554u) | (wv & ~ao & 0xaaaaaaaau)) & (tl - 1) & (test ? ~0u : 0u);
  const uint32_t c = (uint32_t)__popc(d);
  const uint32_t ntop = n + (uint32_t)(31 - __clz((int)d)) - c;  // n + (msb(d) - 2 - (c-1)) + 1: coefficients settled if c > 0
  // data bits two places down (the virtual bit and the first test), then each moved down by its rank:
  // y = R0 - (R1 >> 1) - (R2 >> 2) - (R3 >> 3), Rj = the data bits above the lowest j
  const uint32_t d0 = d >> 2, d1 = d0 & (d0 - 1), d2 = d1 & (d1 - 1), d3 = d2 & (d2 - 1), d4 = d3 & (d3 - 1);
  bool slow = test && (term == 0 || tpos > left || (c != 0 && ntop > N - 1));
  uint32_t y = d0 - (d1 >> 1) - (d2 >> 2) - (d3 >> 3);
  if constexpr (N > 32)
    finish_plane<N>(br, prev);  // the previous plane's leftover work: independent of everything above
  // one vote on the common path: more than four new coefficients, or something the shortcut cannot prove
  bool fast = true;
  if (__any_sync(FULL, slow || d4 != 0)) {
    const uint32_t d5 = d4 & (d4 - 1), d6 = d5 & (d5 - 1), d7 = d6 & (d6 - 1), d8 = d7 & (d7 - 1);
    y -= (d4 >> 4) + (d5 >> 5) + (d6 >> 6) + (d7 >> 7);
    fast = !__any_sync(FULL, slow || d8 != 0);
  }
  if (fast) {
    if constexpr (N > 32)
      rec.ybits = shl64c((uint64_t)y, n);
    else
      rec.ybits = shl32c(y, n);
    const uint32_t used = test ? tpos : 0u;
    bits = left - used;
    br.bp = tp + used;
    n = c ? ntop : n;
  }
  else {
    // exact per-item loop (decode.c:96-117) on this plane for every lane
    R x = 0;
    uint32_t b = left, nn = n, p = tp;
    if (!done) {
      while (b && nn < N) {
        const uint32_t g0 = br.peek32(p);
        b--;
        p++;
        if (!(g0 & 1u))
          break;
        const uint32_t room = N - 1 - nn;
        const uint32_t lim = b < room ? b : room;       // bits the unary scan may read
        uint32_t taken = 0, g = g0 >> 1, avail = 31;
        while (taken < lim) {
          const uint32_t step = lim - taken < avail ? lim - taken : avail;
          const uint32_t gz = g ? (uint32_t)__ffs((int)g) - 1 : 32u;
          if (gz < step) {
            taken += gz + 1;  // gz zeros and the one-bit
            nn += gz;
            break;
          }
          taken += step;
          nn += step;
          g = br.peek32(p + taken);
          avail = 32;
        }
        p += taken;
        b -= taken;
        x |= (R)1 << (nn & (8 * sizeof(R) - 1));  // deposited even if the scan ran dry (nn <= N-1)
        nn++;
      }
    }
    rec.ybits = x;
    bits = b;
    br.bp = p;
    n = nn;
  }
  if constexpr (N <= 32) {
    if (rec.store)
      *rec.dst = verb_small | rec.ybits;
    rec.store = false;
  }
}

// Narrow plane steps (blocks of 64 values), two planes per call: the mirror of encode_pair_narrow.
// Precondition: n <= 32 in every block of the warp.  Verbatim bits and T of a plane come from one
// 64-bit look at the stream; when the plane ends with at most 32 significant coefficients its plane
// word is 32 bits and is complete at once (no deferred verbatim fetch).  The second plane is parsed
// from the first one's tentative results and ONE vote covers both: false - with no state changed and
// nothing stored - when some lane needs the general step for either plane (more than four new
// coefficients, a T the 32-bit window cannot hold, the budget running out inside T, or a coefficient
// beyond the 32nd becoming significant).
struct NarrowParse {
  uint32_t x, bits, bp, n;  // plane word; budget, read position and significant count after the plane
  bool dn, ok;
};
__device__ __forceinline__ NarrowParse narrow_plane_parse(const ColReader& br, bool dn, uint32_t bits, uint32_t bp, uint32_t n)
{
  NarrowParse r;
  const uint32_t m = dn ? 0u : (n < bits ? n : bits);  // verbatim bits (<= 32 when the precondition holds)
  uint32_t lo, hi;
  br.peek64(bp, lo, hi);
  const uint32_t verb = lo & mask32(m);
  const uint32_t w = __funnelshift_rc(lo, hi, m);      // T starts here (m == 32: the high word)
  const uint32_t left = bits - m;                      // budget at T
  const bool test = !dn && left != 0;                  // (n <= 32 < N: there is always a group test)
  // parse T as in decode_plane_lockstep: W' = virtual data bit, then the stream
  const uint32_t wv = (w << 1) | 1u;
  const uint32_t s = wv & ~(wv << 1);                    // run starts
  const uint32_t ae = wv + (s & 0x55555555u), ao = wv + (s & 0xaaaaaaaau);
  const uint32_t term = (ae & ~wv & 0xaaaaaaaau) | (ao & ~wv & 0x55555555u);  // just past each odd-length run
  const uint32_t tl = term & (0u - term);                // lowest end mark; tl - 1 masks the bits of T
  const uint32_t tpos = 31u - (uint32_t)__clz((int)tl);  // bits of T (0xffffffff when no end in the window)
  const uint32_t d = ((wv & ~ae & 0x55555554u) | (wv & ~ao & 0xaaaaaaaau)) & (tl - 1) & (test ? ~0u : 0u);
  const uint32_t c = (uint32_t)__popc(d);
  const uint32_t ntop = n + (uint32_t)(31 - __clz((int)d)) - c;  // coefficients settled if c > 0
  const uint32_t d0 = d >> 2, d1 = d0 & (d0 - 1), d2 = d1 & (d1 - 1), d3 = d2 & (d2 - 1), d4 = d3 & (d3 - 1);
  r.ok = !((test && (term == 0 || tpos > left || (c != 0 && ntop > 32))) || d4 != 0 || n > 32);
  const uint32_t y = d0 - (d1 >> 1) - (d2 >> 2) - (d3 >> 3);  // data bits moved down by 2 + rank
  r.x = verb | shl32c(y, n);                                  // (n == 32 and ok imply y == 0)
  const uint32_t used = test ? tpos : 0u;
  r.dn = dn;
  r.bits = left - used;
  r.bp = bp + m + used;
  r.n = c ? ntop : n;
  return r;
}

template <int N>
__device__ __forceinline__ bool decode_pair_narrow(ColReader& br, int kmin, int k, uint32_t& bits, uint32_t& n, int& lowest,
                                                   bool& done, typename PlaneWord<N>::type* plane1, typename PlaneWord<N>::type* plane2)
{
  const bool dn1 = done || k - 1 < kmin || bits == 0;
  const NarrowParse p1 = narrow_plane_parse(br, dn1, bits, br.bp, n);
  const bool dn2 = dn1 || k - 2 < kmin || p1.bits == 0;
  const NarrowParse p2 = narrow_plane_parse(br, dn2, p1.bits, p1.bp, p1.n);
  if (__any_sync(0xffffffffu, !(p1.ok && p2.ok)))
    return false;
  if (!dn1)
    *plane1 = (typename PlaneWord<N>::type)p1.x;
  if (!dn2)
    *plane2 = (typename PlaneWord<N>::type)p2.x;
  done = dn2;
  lowest = dn2 ? (dn1 ? lowest : k - 1) : k - 2;
  bits = p2.bits;
  br.bp = p2.bp;
  n = p2.n;
  return true;
}

// Small-universe plane steps (blocks of 16 or 64 values), two planes per call: the mirror of encode_planes_small8.
// While at most 8 coefficients are significant and a plane's group-tested part ends within nine bits without
// reaching beyond coefficient 7, the part is ONE look-up in kDecLut8h by (n, the nine bits after the n verbatim
// bits): bits consumed, one-bits deposited, n afterwards.  The second plane is parsed from the first one's
// tentative state and ONE vote covers both: false - nothing stored, no state changed - when some lane needs the
// general step (entry 0 of the table, n > 8, or fewer than n + 9 bits of budget left).
struct Small8Parse {
  uint32_t x, bits, bp, n;
  bool ok;
};
__device__ __forceinline__ Small8Parse small8_plane_parse(const ColReader& br, bool dn, uint32_t bits, uint32_t bp, uint32_t n)
{
  Small8Parse r;
  const uint32_t w = br.peek32(bp);
  const uint32_t nn = n < 8 ? n : 8;                       // (n > 8: not ok, whatever is looked up)
  const uint32_t t = shr32c(w, nn) & 511u;
  uint32_t e;
  asm("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(br.lut8 + (((nn << 9) + t) << 1)));
  r.ok = dn || (n <= 8 && bits >= n + 9 && e != 0);
  const uint32_t used = dn ? 0u : n + (e & 15u);
  r.x = (w & mask32(n)) | ((e >> 4) & 0xffu);
  r.n = dn ? n : e >> 12;
  r.bits = bits - used;
  r.bp = bp + used;
  return r;
}
template <int N>
__device__ __forceinline__ bool decode_pair_small8(ColReader& br, int kmin, int k, uint32_t& bits, uint32_t& n, int& lowest,
                                                   bool& done, typename PlaneWord<N>::type* plane1, typename PlaneWord<N>::type* plane2)
{
  const bool dn1 = done || k - 1 < kmin || bits == 0;
  const Small8Parse p1 = small8_plane_parse(br, dn1, bits, br.bp, n);
  const bool dn2 = dn1 || k - 2 < kmin || p1.bits == 0;
  const Small8Parse p2 = small8_plane_parse(br, dn2, p1.bits, p1.bp, p1.n);
  if (__any_sync(0xffffffffu, !(p1.ok && p2.ok)))
    return false;
  if (!dn1)
    *plane1 = (typename PlaneWord<N>::type)p1.x;
  if (!dn2)
    *plane2 = (typename PlaneWord<N>::type)p2.x;
  done = dn2;
  lowest = dn2 ? (dn1 ? lowest : k - 1) : k - 2;
  bits = p2.bits;
  br.bp = p2.bp;
  n = p2.n;
  return true;
}

template <int N>
__device__ __forceinline__ void decode_planes_lockstep(ColReader& br, int kmin, int klo, int kbase, LockDecodeState& st,
                                                       typename PlaneWord<N>::type* sp)
{
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t bits = st.bits, n = st.n;
  int k = st.k, lowest = st.lowest;
  bool done = st.done;
  PlaneRec<N> rec = { 0, 0, 0, sp, false };
  // blocks of 64 values: narrow steps while no block of the warp has more than 32 significant coefficients
  // (n only grows, so once a pair has gone the general way with n > 32 somewhere the test is skipped)
  bool narrow = N > 32 && ZB_NARROW;
  // small-universe steps until some block of the warp has more than 8 significant coefficients
  bool small8 = (N == 16 || N == 64) && br.lut8 != 0 && !__any_sync(FULL, !done && n > 8);
  while (k > klo && __any_sync(FULL, !done)) {
    if (br.needs_restage()) {  // variable rate, long blocks: slide the window (positions in rec would go stale)
      finish_plane<N>(br, rec);
      rec.store = false;
      br.restage();
    }
    if constexpr (N == 16 || N == 64) {
      if (small8) {
        if constexpr (N > 32) {  // (a general step may have left its verbatim bits for later)
          finish_plane<N>(br, rec);
          rec.store = false;
        }
        if (decode_pair_small8<N>(br, kmin, k, bits, n, lowest, done, sp + (k - 1 - kbase) * 32, sp + (k - 2 - kbase) * 32)) {
          k -= 2;
          continue;
        }
        small8 = !__any_sync(FULL, !done && n > 8);
      }
    }
    if constexpr (N > 32 && ZB_NARROW) {
      if (narrow) {
        if (decode_pair_narrow<N>(br, kmin, k, bits, n, lowest, done, sp + (k - 1 - kbase) * 32, sp + (k - 2 - kbase) * 32)) {
          k -= 2;
          continue;
        }
        narrow = !__any_sync(FULL, n > 32);  // (the general steps below may still find every lane narrow for the next pair)
      }
    }
    decode_plane_lockstep<N>(br, kmin, k - 1, bits, n, lowest, done, sp + (k - 1 - kbase) * 32, rec);
    decode_plane_lockstep<N>(br, kmin, k - 2, bits, n, lowest, done, sp + (k - 2 - kbase) * 32, rec);
    k -= 2;
  }
  finish_plane<N>(br, rec);
  st.bits = bits;
  st.n = n;
  st.k = k;
  st.lowest = lowest;
  st.done = done;
}

// Run/event decoder for blocks of 64 values.  While few coefficients are significant most planes of
// smooth data code as "n verbatim bits, then a lone '0' group test" (about two planes in three on the
// benchmark field), and a run of such planes is a periodic bit pattern: the test bits sit at positions
// n, 2n+1, 3n+2, ... of the stream.  Each lane therefore walks ITS OWN planes (k is per lane here,
// not warp uniform): a step looks at 32 stream bits, finds the first test bit that is set with one
// masked find-first-set against a table of test positions (ColReader::run_mask) - limited to the bits
// the budget still covers and the planes left above the precision limit and the bottom of the resident
// set -, stores the planes before it with a two-instruction extract each, and then parses the one
// plane that follows with the general step (decode_plane_lockstep).  The warp iterates until every
// lane has reached the bottom of the set: about 13 steps instead of 36 plane iterations at rate 8,
// and the empty planes at the top of a block (n = 0) go thirty-two at a time.
template <int N>
__device__ __forceinline__ void decode_planes_events(ColReader& br, int kmin, int klo, int kbase, LockDecodeState& st,
                                                     typename PlaneWord<N>::type* sp)
{
  using PW = typename PlaneWord<N>::type;
  constexpr uint32_t FULL = 0xffffffffu;
  uint32_t bits = st.bits, n = st.n;
  int k = st.k, lowest = st.lowest;
  bool done = st.done;
  PlaneRec<N> rec = { 0, 0, 0, sp, false };
  const int kfloor = klo > kmin ? klo : kmin;  // planes below are not coded (in this set / at all)
  while (__any_sync(FULL, !done && k > klo)) {
    if (br.needs_restage()) {  // variable rate, long blocks: slide the window (positions in rec would go stale)
      finish_plane<N>(br, rec);
      rec.store = false;
      br.restage();
    }
    // (a) the run of planes whose group test fails at once
    {
      const bool active = !done && k > kfloor && n < 32;
      const uint32_t w = br.peek32(br.bp);
      const uint32_t L = n + 1;
      const uint32_t span = (uint32_t)(k - kfloor) * L;               // bits the planes still to come would take
      const uint32_t valid = active ? (br.run_mask(n & 31) & mask32(bits < span ? bits : span)) : 0u;
      const uint32_t hit = w & valid;
      const uint32_t j = (uint32_t)__popc(valid & ((hit & (0u - hit)) - 1u));  // test positions below the first set one
      const uint32_t vm = mask32(n);
      PW* dst = sp + (k - 1 - kbase) * 32;
      uint32_t ww = w;
      for (uint32_t i = 0; i < j; i++) {
        *dst = (PW)(ww & vm);
        dst -= 32;
        ww = shr32c(ww, L);
      }
      br.bp += j * L;
      bits -= j * L;
      k -= (int)j;
      lowest = j ? k : lowest;
    }
    // (b) the plane that follows, by the general step; lanes at the bottom of the set sit it out
    {
      bool d = done || k <= klo;
      decode_plane_lockstep<N>(br, kmin, k - 1, bits, n, lowest, d, sp + (k - 1 - kbase) * 32, rec);
      if (k > klo) {
        done = d;
        k -= d ? 0 : 1;
      }
    }
  }
  finish_plane<N>(br, rec);
  st.bits = bits;
  st.n = n;
  st.k = k;
  st.lowest = lowest;
  st.done = done;
}

// developer switch for A/B timing: 0 keeps the plane-lockstep loop for blocks of 64 values
#ifndef ZB_EVENTS
#define ZB_EVENTS 0
#endif
template <int N>
__device__ __forceinline__ void decode_planes_any(ColReader& br, int kmin, int klo, int kbase, LockDecodeState& st,
                                                  typename PlaneWord<N>::type* sp)
{
  if constexpr (N > 32 && ZB_EVENTS)
    decode_planes_events<N>(br, kmin, klo, kbase, st, sp);
  else
    decode_planes_lockstep<N>(br, kmin, klo, kbase, st, sp);
}

// ------------------------------------------------------------------------------------------------
// floating-point helpers
// ------------------------------------------------------------------------------------------------
template <class T> struct FpBits;
template <> struct FpBits<float> {
  using U = uint32_t;
  __device__ static __forceinline__ U bits(float f) { return __float_as_uint(f); }
  __device__ static __forceinline__ float make(U b) { return __uint_as_float(b); }
};
template <> struct FpBits<double> {
  using U = uint64_t;
  __device__ static __forceinline__ U bits(double f) { return (U)__double_as_longlong(f); }
  __device__ static __forceinline__ double make(U b) { return __longlong_as_double((long long)b); }
};

// frexp exponent of the largest finite magnitude in the block, clamped as encodef.c:10-27;
// NaNs never win the comparison (encodef.c:35-37 uses `max < f`).  EXACT = false (lossy modes,
// where NaN / infinity inputs are undefined behaviour upstream, docs/source/faq.rst:282-289) takes
// the maximum over the words holding the exponent only and ORs the rest for the zero test.
template <class TR, int N, bool EXACT>
__device__ __forceinline__ int block_emax(const typename TR::Scalar (&v)[N])
{
  using U = typename TR::UInt;
  const U absmask = ~(U)0 >> 1, infbits = (U)((1u << TR::EBITS) - 1) << TR::MANT;
  if constexpr (!EXACT && sizeof(U) == 8) {
    // The high words are sign-magnitude: as signed integers the maximum is the largest positive
    // value (or, with no positive value, the negative one of largest magnitude), as unsigned
    // integers it is the negative value of largest magnitude (or the largest positive one); the
    // larger of the two magnitudes is the block maximum.  Two 3-input max chains, no masking.
    int32_t smax = (int32_t)0x80000000;
    uint32_t umax = 0;
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const uint32_t h0 = (uint32_t)(FpBits<typename TR::Scalar>::bits(v[i]) >> 32);
      const uint32_t h1 = (uint32_t)(FpBits<typename TR::Scalar>::bits(v[(i + 1) % N]) >> 32);
      smax = __vimax3_s32(smax, (int32_t)h0, (int32_t)h1);
      umax = __vimax3_u32(umax, h0, h1);
    }
    const uint32_t a = (uint32_t)smax & 0x7fffffffu, b = umax & 0x7fffffffu;
    const uint32_t top = a > b ? a : b;
    const int E = (int)(top >> 20);
    if (E) return E - TR::EBIAS + 1;
    // subnormal or zero block (rare): look at all the bits
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      const uint64_t w = FpBits<typename TR::Scalar>::bits(v[i]);
      any |= ((uint32_t)(w >> 32) & 0x7fffffffu) | (uint32_t)w;
    }
    return any ? 1 - TR::EBIAS : -TR::EBIAS;
  }
  U m = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    U a = FpBits<typename TR::Scalar>::bits(v[i]) & absmask;
    if (EXACT) a = a > infbits ? 0 : a;
    m = a > m ? a : m;
  }
  int E = (int)(m >> TR::MANT);
  // inf: frexp leaves the exponent 0 in glibc; the value is irrelevant but keep it
  if (EXACT && m == infbits) return 0 > 1 - TR::EBIAS ? 0 : 1 - TR::EBIAS;
  if (E) return E - TR::EBIAS + 1;
  return m ? 1 - TR::EBIAS : -TR::EBIAS;
}

// 2^e as Scalar for e in the normal range, +inf above it, 0 / subnormal below it
template <class Scalar> __device__ __forceinline__ Scalar pow2(int e);
template <> __device__ __forceinline__ float pow2<float>(int e)
{
  if (e > 127) return __uint_as_float(0x7f800000u);
  if (e >= -126) return __uint_as_float((uint32_t)(e + 127) << 23);
  return e >= -149 ? __uint_as_float(1u << (e + 149)) : 0.0f;
}
template <> __device__ __forceinline__ double pow2<double>(int e)
{
  if (e > 1023) return __longlong_as_double(0x7ff0000000000000ll);
  if (e >= -1022) return __longlong_as_double((long long)(e + 1023) << 52);
  return e >= -1074 ? __longlong_as_double(1ll << (e + 1074)) : 0.0;
}

__device__ __forceinline__ int32_t cvt_rz(float p) { return __float2int_rz(p); }
__device__ __forceinline__ int64_t cvt_rz(double p) { return __double2ll_rz(p); }
__device__ __forceinline__ float cvt_rn(int32_t i, float) { return __int2float_rn(i); }
__device__ __forceinline__ double cvt_rn(int64_t i, double) { return __ll2double_rn(i); }

// (Int)(2^(P-2-emax) * f), truncating; an infinite scale reproduces the x86 "integer
// indefinite" result of the reference build (oracle/zfp_oracle.c cast_fwd)
template <class TR, int N>
__device__ __forceinline__ void cast_fwd(typename TR::Int (&q)[N], const typename TR::Scalar (&v)[N], int emax)
{
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  const int se = TR::P - 2 - emax;
  const Scalar s = pow2<Scalar>(se);
  const bool overflow = se > (TR::P == 32 ? 127 : 1023);
  // tiny block maxima (infinite scale) are rare: whole warps skip the per-value selects
  if (!__any_sync(__activemask(), overflow)) {
#pragma unroll
    for (int i = 0; i < N; i++)
      q[i] = cvt_rz(s * v[i]);
    return;
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    Int r = cvt_rz(s * v[i]);
    q[i] = overflow ? (Int)((typename TR::UInt)1 << (TR::P - 1)) : r;
  }
}

// (Scalar)i * 2^(emax-(P-2))
template <class TR, int N>
__device__ __forceinline__ void cast_inv(typename TR::Scalar (&v)[N], const typename TR::Int (&q)[N], int emax)
{
  using Scalar = typename TR::Scalar;
  const Scalar s = pow2<Scalar>(emax - (TR::P - 2));
#pragma unroll
  for (int i = 0; i < N; i++)
    v[i] = s * cvt_rn(q[i], Scalar());
}

template <class TR>
__device__ __forceinline__ uint32_t block_precision(int emax, uint32_t maxprec, int minexp, int dims)
{
  int p = emax - minexp + 2 * dims + 2;
  p = p < 0 ? 0 : p;
  return (uint32_t)p < maxprec ? (uint32_t)p : maxprec;
}

// negabinary
template <class Int>
__device__ __forceinline__ typename std::make_unsigned<Int>::type int2uint(Int x)
{
  using U = typename std::make_unsigned<Int>::type;
  const U mask = (U)0xaaaaaaaaaaaaaaaaull;
  return ((U)x + mask) ^ mask;
}
template <class UInt>
__device__ __forceinline__ typename std::make_signed<UInt>::type uint2int(UInt u)
{
  const UInt mask = (UInt)0xaaaaaaaaaaaaaaaaull;
  return (typename std::make_signed<UInt>::type)((u ^ mask) - mask);
}

template <class W, class = void> struct is_lockstep : std::false_type {};
template <class W> struct is_lockstep<W, std::enable_if_t<W::kLockstep>> : std::true_type {};

// ------------------------------------------------------------------------------------------------
// whole-block encode / decode for one thread.  `sp` is the lane's plane column in shared memory.
// Returns the number of bits the block occupies in the stream.
// ------------------------------------------------------------------------------------------------
// SYNC_THREADS > 0 (developer experiment, fixed-rate staged kernel with large CTAs only): the warps that share a
// scheduler (warp, warp + 4, ...; SYNC_THREADS of them in threads) meet at a named barrier before each long
// straight-line stage, so that they walk it together and one instruction fetch serves them all
template <int SYNC_THREADS>
__device__ __forceinline__ void stage_rendezvous()
{
  if constexpr (SYNC_THREADS > 0)
    asm volatile("bar.sync %0, %1;" ::"r"(1 + ((threadIdx.x >> 5) & 3)), "n"(SYNC_THREADS) : "memory");
}

template <int TYPE, int DIMS, bool REV, class Writer, int SYNC_THREADS = 0>
__device__ __forceinline__ uint32_t encode_block(const typename Traits<TYPE>::Scalar (&v)[1 << (2 * DIMS)],
                                                 const Params& prm, Writer& bw,
                                                 typename PlaneWord<(1 << (2 * DIMS))>::type* sp)
{
  using TR = Traits<TYPE>;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int N = 1 << (2 * DIMS), P = TR::P;
  constexpr bool reversible = REV;  // prm.minexp < ZFP_MIN_EXP, resolved by the launcher
  uint32_t bits = 0, maxprec = prm.maxprec;
  // A block that codes as a single '0' bit skips the coefficient stage by flag, not by an early
  // return: in the lockstep kernels all 32 lanes of the warp must reach the warp votes below.
  bool coded = true, pad = true;
  Int q[N];

  if constexpr (TR::is_fp) {
    const int emax = block_emax<TR, N, REV>(v);
    if (!reversible) {
      maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, DIMS);
      const uint32_t e = maxprec ? (uint32_t)(emax + TR::EBIAS) : 0;
      coded = e != 0;
      bits = coded ? 1 + TR::EBITS : 1;
      bw.put(coded ? 2 * (uint64_t)e + 1 : 0, bits);
      cast_fwd<TR>(q, v, emax);
    }
    else {
      // try the block-floating-point route and verify it bit for bit (revencodef.c:6-27)
      bool same = true;
      if (emax != -TR::EBIAS) {
        typename TR::Scalar back[N];
        cast_fwd<TR>(q, v, emax);
        cast_inv<TR>(back, q, emax);
#pragma unroll
        for (int i = 0; i < N; i++)
          same &= FpBits<typename TR::Scalar>::bits(back[i]) == FpBits<typename TR::Scalar>::bits(v[i]);
      }
      else {
#pragma unroll
        for (int i = 0; i < N; i++) {
          q[i] = 0;
          same &= FpBits<typename TR::Scalar>::bits(v[i]) == 0;
        }
      }
      if (same) {
        const uint32_t e = (uint32_t)(emax + TR::EBIAS);
        if (!e) {
          bw.put(0, 1);
          bits = 1;
          coded = pad = false;  // no minbits padding on this path (revencodef.c:64-69)
        }
        else {
          bw.put(1, 2);
          bw.put(e, TR::EBITS);
          bits = 2 + TR::EBITS;
        }
      }
      else {
        // sign-magnitude bit patterns -> two's complement (revencodef.c:29-41)
#pragma unroll
        for (int i = 0; i < N; i++) {
          Int x = (Int)FpBits<typename TR::Scalar>::bits(v[i]);
          q[i] = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        }
        bw.put(3, 2);
        bits = 2;
      }
    }
  }
  else {
#pragma unroll
    for (int i = 0; i < N; i++)
      q[i] = (Int)v[i];
  }

  UInt u[N];
  UInt any = 0;  // reversible mode: OR of all coefficients
  // lossy modes: u holds q + 0xaaaa..., the XOR half of int2uint happens inside the plane transposes
  constexpr int NEG = REV ? 0 : 1;
  if (!reversible) {
    stage_rendezvous<SYNC_THREADS>();
    xform_fwd<0, DIMS>(q);
#pragma unroll
    for (int i = 0; i < N; i++)
      u[i] = (UInt)q[perm_at<DIMS>(i)] + (UInt)0xaaaaaaaaaaaaaaaaull;
  }
  else {
    xform_fwd<2, DIMS>(q);
#pragma unroll
    for (int i = 0; i < N; i++) {
      u[i] = int2uint(q[perm_at<DIMS>(i)]);
      any |= u[i];
    }
    // precision = width - (trailing zeros common to all coefficients), in [1, maxprec]
    uint32_t prec = any ? (uint32_t)P - (P == 64 ? ctz64((uint64_t)any) : (uint32_t)__ffs((int)any) - 1) : 0;
    prec = prec < prm.maxprec ? prec : prm.maxprec;
    prec = prec > 1 ? prec : 1;
    if (coded) {
      bw.put(prec - 1, TR::PBITS);
      bits += TR::PBITS;
    }
    maxprec = prec;
  }
  if constexpr (is_lockstep<Writer>::value) {
    // plane-lockstep coder (column writer), planes made resident progressively
    const uint32_t budget = prm.maxbits - bits, start = bw.tell();
    const uint32_t limit = start + budget < start ? 0xffffffffu : start + budget;  // write position where the budget ends
    const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
    LockState st = { 0, P, !coded };
    if constexpr (REV) {
      // Reversible residuals leave many high planes empty.  With no coefficient significant yet an
      // empty plane codes as a lone '0' test, so the planes above the warp's highest occupied plane
      // go out in one append and the lockstep walk starts below them (two planes per step: even k).
      const int top = any ? (P == 64 ? 63 - __clzll((long long)any) : 31 - __clz((int)any)) : -1;
      int first = coded ? top + 1 : 0;                 // planes >= first are empty in this block
      if (coded && budget < (uint32_t)P) first = P;    // the budget could bind: no shortcut
      int kstart = (int)__reduce_max_sync(0xffffffffu, (unsigned)first);
      kstart = (kstart + 1) & ~1;
      if (kstart < P) {
        const int lo = kstart > kmin ? kstart : kmin;  // this block codes planes P-1 .. lo as '0'
        bw.append64(0, 0, (coded && lo < P) ? (uint32_t)(P - lo) : 0u);
        st.k = kstart;
      }
    }
    if constexpr (P == 64 && N == 64) {
      // the high 32 planes as one half, then the low half as two windows of 16 planes, each only if
      // some block of the warp still has planes and budget left when it gets there (all 32 lanes
      // reach the votes: the lockstep kernels have no early exit).  Measured on 1024^3 fp64:
      // blocks that stop within the high half (accuracy 1e-6, precision 32) are fastest with an
      // undivided half, blocks that go a few planes further (rate 8) with a 16-plane window.
      if (st.k > 32) {
        stage_rendezvous<SYNC_THREADS>();
        int ksmall = 32;
        const int kup = to_planes_half<1, NEG, UInt, N>(u, sp, (!REV && bw.lut) ? &ksmall : nullptr);
        if constexpr (!REV)
          if (bw.lut)
            encode_planes_small8<N>(bw, limit, kmin, 32, ksmall, st, sp);
        encode_planes_lockstep<N>(bw, limit, kmin, 32, 32, st, sp, 32 + kup);
      }
#pragma unroll 1
      for (int w = 1; w >= 0; w--) {
        if (!__any_sync(0xffffffffu, !st.done && st.k > kmin && bw.tell() < limit))
          break;
        if (st.k <= 16 * w)
          continue;
        const int kup = to_planes_window<0, NEG>(u, sp, (uint32_t)w);
        encode_planes_lockstep<N>(bw, limit, kmin, 16 * w, 16 * w, st, sp, 16 * w + kup);
      }
    }
    else if constexpr (P == 64) {
      if (st.k > 32) {
        int ksmall = 32;
        const int kup = to_planes_half<1, NEG, UInt, N>(u, sp, (N == 16 && !REV && bw.lut) ? &ksmall : nullptr);
        if constexpr (N == 16 && !REV)
          if (bw.lut)
            encode_planes_small8<N>(bw, limit, kmin, 32, ksmall, st, sp);
        encode_planes_lockstep<N>(bw, limit, kmin, 32, 32, st, sp, 32 + kup);
      }
      if (__any_sync(0xffffffffu, !st.done && st.k > kmin && bw.tell() < limit)) {
        const int kup = to_planes_half<0, NEG, UInt, N>(u, sp);
        encode_planes_lockstep<N>(bw, limit, kmin, 0, 0, st, sp, kup);
      }
    }
    else {
      int ksmall = 32;
      const int kup = to_planes_half<0, NEG, UInt, N>(u, sp, ((N == 16 || N == 64) && !REV && bw.lut) ? &ksmall : nullptr);
      if constexpr ((N == 16 || N == 64) && !REV)
        if (bw.lut)
          encode_planes_small8<N>(bw, limit, kmin, 0, ksmall, st, sp);
      encode_planes_lockstep<N>(bw, limit, kmin, 0, 0, st, sp, kup);
    }
    const uint32_t used = bw.tell() - start;
    bits += used < budget ? used : budget;
  }
  else if (coded) {
    to_planes<NEG, UInt, N>(u, sp);
    bits += encode_planes<N, P>(bw, prm.maxbits - bits, maxprec, sp);
  }
  if (pad && bits < prm.minbits) {
    bw.pad(prm.minbits - bits);
    bits = prm.minbits;
  }
  return bits;
}

// Register budget of a warpgroup (PTX setmaxnreg; all four warps of the group execute it together)
template <int REGS> __device__ __forceinline__ void wg_reg_release() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void wg_reg_acquire() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
// the warps of a group meet (named barrier 1 + group index); a group is one warpgroup or, with
// ZB_PS_GROUP_THREADS = 256, two that change their budget together
#ifndef ZB_PS_GROUP_THREADS
#define ZB_PS_GROUP_THREADS 128
#endif
__device__ __forceinline__ void wg_barrier()
{
  asm volatile("bar.sync %0, %1;" ::"r"(1 + threadIdx.x / ZB_PS_GROUP_THREADS), "n"(ZB_PS_GROUP_THREADS) : "memory");
}

// The large register budget goes to whole warpgroups: setmaxnreg allocates warp by warp, and warps of three groups
// that each hold a part of the pool would wait for one another at their group barriers for ever.  A counter in
// shared memory (set by the kernel to the number of groups the pool can make large) is taken by a group's first
// thread before any of its warps asks for registers, so the hardware request never has to wait.
__device__ __forceinline__ int* wg_large_slots()
{
  __shared__ int slots;
  return &slots;
}
template <int REGS>
__device__ __forceinline__ void wg_enter_large()
{
  wg_barrier();
  if (threadIdx.x % ZB_PS_GROUP_THREADS == 0) {
    int* slots = wg_large_slots();
    while (atomicSub(slots, 1) <= 0) {
      atomicAdd(slots, 1);
      __nanosleep(200);
    }
  }
  wg_barrier();
  wg_reg_acquire<REGS>();
}
template <int REGS>
__device__ __forceinline__ void wg_leave_large()
{
  wg_reg_release<REGS>();
  wg_barrier();
  if (threadIdx.x % ZB_PS_GROUP_THREADS == 0)
    atomicAdd(wg_large_slots(), 1);
}

// PS > 0 (kernels_ps.cuh, blocks of 64 64-bit values): the caller's warpgroup holds a small register
// budget while it parses planes 63..32 and acquires PS registers per thread before the first transposes.
// LENGTH_ONLY (serial readers): parse the block for its length and skip transposes, transform and cast (index rebuild).
template <int TYPE, int DIMS, bool REV, class Reader, int PS = 0, bool LENGTH_ONLY = false>
__device__ __forceinline__ uint32_t decode_block(typename Traits<TYPE>::Scalar (&v)[1 << (2 * DIMS)],
                                                 const Params& prm, Reader& br,
                                                 typename PlaneWord<(1 << (2 * DIMS))>::type* sp)
{
  using TR = Traits<TYPE>;
  using Scalar = typename TR::Scalar;
  using Int = typename TR::Int;
  using UInt = typename TR::UInt;
  constexpr int N = 1 << (2 * DIMS), P = TR::P;
  constexpr bool reversible = REV;  // prm.minexp < ZFP_MIN_EXP, resolved by the launcher
  uint32_t bits = 0, maxprec = prm.maxprec;
  int emax = 0;
  bool reinterpret = false;
  bool zero = false;  // an all-zero block (single '0' bit) flows through by flag: see encode_block

  if constexpr (TR::is_fp) {
    bits = 1;
    zero = !br.get(1);
    if (!zero) {
      if (!reversible) {
        bits += TR::EBITS;
        emax = (int)br.get(TR::EBITS) - TR::EBIAS;
        maxprec = block_precision<TR>(emax, prm.maxprec, prm.minexp, DIMS);
      }
      else {
        bits++;
        reinterpret = br.get(1) != 0;
        if (!reinterpret) {
          bits += TR::EBITS;
          emax = (int)br.get(TR::EBITS) - TR::EBIAS;
        }
      }
    }
  }
  if (reversible && !zero) {
    maxprec = (uint32_t)br.get(TR::PBITS) + 1;
    bits += TR::PBITS;
  }

  UInt u[N];
  int lowest = 0;  // lowest plane this block decoded (lockstep readers; P when none)
  // lossy modes: the plane transposes deliver u ^ 0xaaaa... (the XOR half of uint2int); planes that
  // are never decoded leave the mask's bits
  constexpr int NEG = REV ? 0 : 2;
  if constexpr (is_lockstep<Reader>::value) {
    const uint32_t budget = prm.maxbits - bits;
    const int kmin = P > (int)maxprec ? P - (int)maxprec : 0;
    LockDecodeState st = { budget, 0, P, P, zero };
    if constexpr (!PS) {
#pragma unroll
      for (int i = 0; i < N; i++)
        u[i] = (UInt)NegaWord<NEG>::w64;
    }
    if constexpr (REV) {
      // mirror of the encoder's shortcut: leading '0' tests while no coefficient is significant are
      // empty planes; the warp skips the ones all its blocks have in common (even count, at most the
      // 32 planes of the first window, stored as zeros)
      const uint32_t w0 = br.peek32(br.bp);
      const uint32_t lim = st.bits < (uint32_t)(P - kmin) ? st.bits : (uint32_t)(P - kmin);
      uint32_t z = w0 ? (uint32_t)__ffs((int)w0) - 1 : 32u;
      z = zero ? 32u : (z < lim ? z : lim);
      const uint32_t skip = __reduce_min_sync(0xffffffffu, z) & ~1u;
      for (uint32_t j = 0; j < skip; j++)
        sp[(31 - j) * 32] = 0;
      if (!zero) {
        br.bp += skip;
        st.bits -= skip;
      }
      st.k = P - (int)skip;
    }
    if constexpr (P == 64 && N == 64) {
      // (coefficients 32..63 are transposed only once some block of the warp has one of them significant)
      decode_planes_any<N>(br, kmin, 32, 32, st, sp);
      if constexpr (PS) {
        wg_enter_large<PS>();
#pragma unroll
        for (int i = 0; i < N; i++)
          u[i] = (UInt)NegaWord<NEG>::w64;
      }
      from_planes_half<1, NEG, UInt, N>(u, sp, st.lowest, __any_sync(0xffffffffu, st.n > 32));
      if (__any_sync(0xffffffffu, !st.done && st.k > kmin && st.bits != 0)) {
        decode_planes_any<N>(br, kmin, 16, 16, st, sp);
        from_planes_window<1, NEG>(u, sp, st.lowest, __any_sync(0xffffffffu, st.n > 32));
        if (__any_sync(0xffffffffu, !st.done && st.k > kmin && st.bits != 0)) {
          decode_planes_any<N>(br, kmin, 0, 0, st, sp);
          from_planes_window<0, NEG>(u, sp, st.lowest, __any_sync(0xffffffffu, st.n > 32));
        }
      }
      // the warps of the CTA (PS: of the warpgroup) enter the long straight-line tail together (shared instruction fetch)
      if constexpr (PS)
        wg_barrier();
      else if constexpr (ZB_DEC_TAIL_SYNC || REV)  // (reversible mode keeps the long integer tail, and the rendezvous: 13.4 against 13.7 ms)
        __syncthreads();
    }
    else if constexpr (P == 64) {
      decode_planes_lockstep<N>(br, kmin, 32, 32, st, sp);
      from_planes_half<1, NEG, UInt, N>(u, sp, st.lowest);
      if (__any_sync(0xffffffffu, !st.done && st.k > kmin && st.bits != 0)) {
        decode_planes_lockstep<N>(br, kmin, 0, 0, st, sp);
        from_planes_half<0, NEG, UInt, N>(u, sp, st.lowest);
      }
      __syncthreads();
    }
    else {
      decode_planes_any<N>(br, kmin, 0, 0, st, sp);
      from_planes_half<0, NEG, UInt, N>(u, sp, st.lowest, N <= 32 || __any_sync(0xffffffffu, st.n > 32));
    }
    bits += budget - st.bits;
    lowest = st.lowest;
  }
  else if (!zero) {
    bool direct = false;
    if constexpr (N == 4 && !is_lockstep<Reader>::value) {
      if (br.lut4) {
        bits += decode_planes4_direct<P, NEG>(br, prm.maxbits - bits, maxprec, u);
        direct = true;
      }
    }
    if (!direct) {
      int kstop;
      bits += decode_planes<N, P, Reader, !LENGTH_ONLY>(br, prm.maxbits - bits, maxprec, sp, kstop);  // (length only: sp unused)
      if constexpr (!LENGTH_ONLY)
        from_planes<NEG, UInt, N>(u, sp, kstop);
    }
  }
  else {
#pragma unroll
    for (int i = 0; i < N; i++)
      u[i] = (UInt)NegaWord<NEG>::w64;
  }
  if (zero)
    bits = 1;
  if (bits < prm.minbits)
    bits = prm.minbits;
  if constexpr (LENGTH_ONLY)
    return bits;

  Int q[N];
#pragma unroll
  for (int i = 0; i < N; i++)
    q[perm_at<DIMS>(i)] = REV ? uint2int(u[i]) : (Int)(u[i] - (UInt)0xaaaaaaaaaaaaaaaaull);

#if ZB_FP64_TAIL
  if constexpr (TYPE == T_DOUBLE && !REV && DIMS <= 3 && is_lockstep<Reader>::value) {
    // FP64-pipe tail (see inv_lift_f64): every block of the warp stopped at plane >= 10 + 2 DIMS, and
    // the weighted sum of magnitudes proves that no shifted or returned value leaves the int64 range
    constexpr uint32_t FULL = 0xffffffffu;
    if (__all_sync(FULL, lowest >= 10 + 2 * DIMS)) {
      double d[N];
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int i = 0; i < N; i++)
        d[i] = __ll2double_rn(q[i]);  // exact: a multiple of 2^lowest below 2^63
#pragma unroll
      for (int i = 0; i < N; i += 4) {
        s0 = __fma_rn(fabs(d[i]), lift_gain(i, DIMS), s0);
        s1 = __fma_rn(fabs(d[i + 1]), lift_gain(i + 1, DIMS), s1);
        s2 = __fma_rn(fabs(d[i + 2]), lift_gain(i + 2, DIMS), s2);
        s3 = __fma_rn(fabs(d[i + 3]), lift_gain(i + 3, DIMS), s3);
      }
      if (__all_sync(FULL, (s0 + s1) + (s2 + s3) < 0x1.fffffffp+62)) {
        xform_inv_f64<DIMS>(d);
        const double s = pow2<double>(emax - (TR::P - 2));
#pragma unroll
        for (int i = 0; i < N; i++)
          v[i] = s * d[i];
        return bits;
      }
      // back to integers for the general tail (exact both ways; q and d never live together)
#pragma unroll
      for (int i = 0; i < N; i++)
        q[i] = __double2ll_rn(d[i]);
    }
  }
#endif
  if (!reversible)
    xform_inv<1, DIMS>(q);
  else
    xform_inv<3, DIMS>(q);

  if constexpr (TR::is_fp) {
    if (reinterpret) {
#pragma unroll
      for (int i = 0; i < N; i++) {
        Int x = q[i];
        x = x < 0 ? (Int)((UInt)x ^ (~(UInt)0 >> 1)) : x;
        v[i] = FpBits<Scalar>::make((typename FpBits<Scalar>::U)x);
      }
    }
    else if (reversible && emax == -TR::EBIAS) {
#pragma unroll
      for (int i = 0; i < N; i++)
        v[i] = (Scalar)0;
    }
    else
      cast_inv<TR>(v, q, emax);
  }
  else {
#pragma unroll
    for (int i = 0; i < N; i++)
      v[i] = (Scalar)q[i];
  }
  return bits;
}

}  // namespace zb
